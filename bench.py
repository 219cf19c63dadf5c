#!/usr/bin/env python
"""Headline benchmark of the matching hot path (BASELINE.json): descriptor-pairs/s of brute-force SIFT kNN
(k=2) + Lowe ratio test, 1 M queries x 1 M database rows per GPU (config "synthetic SIFT 128-d float, 1M
query x 1M database kNN k=2 ratio test on 1 B200").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (libhpusm.so through the C ABI)
    python bench.py --impl reference ...                      # the reference's CPU path (cv2.BFMatcher + ratio loop)

One step = one full kNN-2 + ratio pass of the query set over this rank's database shard.  With N > 1 the database
is sharded by rows across ranks (weak scaling: every rank owns `--nt` rows), queries are replicated, each rank
produces a local top-2, and one NCCL all-gather of the 16 B/query partial results followed by the lexicographic
merge + ratio kernel gives every rank the global answer.  Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "humanpose-and-urbanscene-matching_b200"
FLOP_PER_PAIR = 256.0  # 2*D, D = 128: norms, top-2 and re-rank are not counted


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nq", type=int, default=1 << 20)
    ap.add_argument("--nt", type=int, default=1 << 20, help="database rows PER GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def sift_like_torch(n, seed, device, d=128):
    """Same recipe as synth.sift_like, generated on the device (1 M x 128 in numpy costs seconds per array)."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch._standard_gamma(torch.full((n, d), 0.6, device=device), generator=g)
    x = x / (x.norm(dim=1, keepdim=True) + 1e-12)
    x = torch.clamp(x, max=0.2)
    x = x / (x.norm(dim=1, keepdim=True) + 1e-12)
    return torch.clamp(torch.round(512.0 * x), max=255.0).float().contiguous()


def make_workload(nq, nt, rank, device, planted=0.1):
    t = sift_like_torch(nt, 1000 + rank, device)
    q = sift_like_torch(nq, 7, device)
    k = int(planted * nq) // max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if k:
        g = torch.Generator(device=device).manual_seed(99 + rank)
        qi = torch.randperm(nq, generator=g, device=device)[:k]
        ti = torch.randint(0, nt, (k,), generator=g, device=device)
        noise = torch.round(torch.randn((k, 128), generator=g, device=device) * 4.0)
        q[qi] = torch.clamp(t[ti] + noise, 0, 255)
    return q, t


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (profiling recipe's clocks line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- reference CPU path (cv2.BFMatcher.knnMatch + the ratio loop of features.filter_matches) --------------------
CV2_MAX_TRAIN = 1 << 17  # cv2.BFMatcher asserts train rows < 2**18 (IMGIDX_ONE): sweep the database in chunks


def reference_step(q_np, t_np, ratio=0.8):
    """features.py:59 + :45-55 without the keypoint gather: returns the number of ratio survivors.  The database is
    swept in chunks below cv2's 2**18-row limit and the per-chunk top-2 lists are merged (lowest index wins ties)."""
    import cv2
    matcher = cv2.BFMatcher(cv2.NORM_L2)
    nq = len(q_np)
    if len(t_np) <= CV2_MAX_TRAIN:
        raw = matcher.knnMatch(q_np, trainDescriptors=t_np, k=2)
        return sum(1 for m in raw if len(m) == 2 and m[0].distance < m[1].distance * ratio)
    best_d = np.full((nq, 2), np.inf, np.float32)
    for lo in range(0, len(t_np), CV2_MAX_TRAIN):
        raw = matcher.knnMatch(q_np, trainDescriptors=t_np[lo:lo + CV2_MAX_TRAIN], k=2)
        d = np.array([[m[0].distance, m[1].distance if len(m) > 1 else np.inf] for m in raw], np.float32)
        best_d = np.sort(np.concatenate([best_d, d], 1), axis=1, kind="stable")[:, :2]
    kept = 0
    for d0, d1 in best_d.tolist():  # the ratio test stays a Python-double comparison as in features.py:48
        if d1 != float("inf") and d0 < d1 * ratio:
            kept += 1
    return kept


def cpu_threads():
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    return int(cv2.getNumThreads())


def pick_cpu_sample(nt, want):
    if want:
        return want
    return int(max(64, min(4096, 2.5e9 // max(nt, 1))))  # about 10-20 s of host work at 1-3e8 pairs/s


def run_reference(args, rank):
    if rank != 0:
        return None
    cores = cpu_threads()
    hp = importlib.import_module(PKG)
    ns = pick_cpu_sample(args.nt, args.cpu_sample)
    q, t = hp.synth.sift_like(ns, args.nt, seed=7, planted=0.1)
    for _ in range(min(args.warmup, 1)):
        reference_step(q[:max(16, ns // 8)], t)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        reference_step(q, t)
        times.append(time.perf_counter() - t0)
    total = float(sum(times))
    value = ns * args.nt * args.steps / total
    return {
        "metric": "descriptor-pairs/sec (kNN+ratio)", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "synthetic SIFT 128-d float, 1M query x 1M database kNN k=2 ratio test on 1 B200",
                   "nq": args.nq, "nt_per_gpu": args.nt, "d": 128, "ratio": 0.8},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference",
                         "sample": "%d queries x %d db rows per step, cv2.BFMatcher(NORM_L2).knnMatch k=2 + ratio loop "
                                   "(features.py:59,:45-55)" % (ns, args.nt)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ---- our path ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, device):
    import torch.distributed as dist
    hp = importlib.import_module(PKG)
    eng = hp.engine
    nq, nt = args.nq, args.nt
    q, t = make_workload(nq, nt, rank, device)
    ws = eng.knn2_l2_workspace(nq, nt, 128, eng.L2_AUTO, device)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=device)
    dst = torch.empty((nq, 2), dtype=torch.float32, device=device)
    gathered_i = torch.empty((world, nq, 2), dtype=torch.int32, device=device) if world > 1 else None
    gathered_d = torch.empty((world, nq, 2), dtype=torch.float32, device=device) if world > 1 else None

    def step():
        eng.knn2_l2(q, t, mode=eng.L2_AUTO, out=(idx, dst), workspace=ws)
        if world > 1:
            gi = torch.where(idx >= 0, idx + rank * nt, idx)
            dist.all_gather_into_tensor(gathered_i, gi)
            dist.all_gather_into_tensor(gathered_d, dst)
            mi, md = eng.knn2_merge(gathered_i, gathered_d)
            return eng.ratio_filter(mi, md, 0.8)
        return eng.ratio_filter(idx, dst, 0.8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
    eng.profile_enable(True)
    eng.profile_collect()
    launches0 = hp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        keep, count = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms, kernel_n = eng.profile_collect()
    eng.profile_enable(False)
    launches = hp.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    survivors = int(count.item())
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())

    # end to end through the public API with HOST buffers: H2D of the descriptors and D2H of the answer every step
    hq, ht = q.cpu().pin_memory(), t.cpu().pin_memory()
    h_idx = torch.empty((nq, 2), dtype=torch.int32).pin_memory()
    h_dst = torch.empty((nq, 2), dtype=torch.float32).pin_memory()
    h_keep = torch.empty((nq,), dtype=torch.uint8).pin_memory()
    dq, dt = torch.empty_like(q), torch.empty_like(t)

    def e2e_step():
        dq.copy_(hq, non_blocking=True)
        dt.copy_(ht, non_blocking=True)
        eng.knn2_l2(dq, dt, mode=eng.L2_AUTO, out=(idx, dst), workspace=ws)
        if world > 1:
            gi = torch.where(idx >= 0, idx + rank * nt, idx)
            dist.all_gather_into_tensor(gathered_i, gi)
            dist.all_gather_into_tensor(gathered_d, dst)
            mi, md = eng.knn2_merge(gathered_i, gathered_d)
        else:
            mi, md = idx, dst
        kp, _ = eng.ratio_filter(mi, md, 0.8)
        h_idx.copy_(mi, non_blocking=True)
        h_dst.copy_(md, non_blocking=True)
        h_keep.copy_(kp, non_blocking=True)

    e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    if world > 1:
        tms = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms.item())
    assert int(h_keep.sum()) == survivors, "e2e pass disagrees with the resident pass"

    if rank != 0:
        return None
    pairs_per_step = float(nq) * float(nt) * world
    value = pairs_per_step * args.steps / (ms * 1e-3)
    pk = peaks()
    peak_tf = (pk or {}).get("bf16_tflops_sustained", 1400.0)
    k_ms = kernel_ms / max(kernel_n, 1)
    flops_per_launch = FLOP_PER_PAIR * float(nq) * float(nt) * args.steps / max(kernel_n, 1)
    achieved = flops_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    mode = {0: "auto", 1: "simt-fp32", 2: "tcgen05-bf16-exact-int", 3: "tcgen05-bf16-split"}.get(eng.knn2_l2_last_mode(), "?")
    out = {
        "metric": "descriptor-pairs/sec (kNN+ratio)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 operands (exact for integer-valued SIFT) / f32 accumulate + f32 re-rank"
        if mode.startswith("tcgen05") else "f32", "data": "synthetic",
        "config": {"workload": "synthetic SIFT 128-d float, 1M query x 1M database kNN k=2 ratio test on 1 B200",
                   "nq": nq, "nt_per_gpu": nt, "d": 128, "ratio": 0.8, "engine": mode, "sharding": "database rows",
                   "l2_flush": "inputs (2 x %.0f MB) exceed the 126 MB L2" % (nq * 512 / 1e6), "ratio_survivors": survivors},
        "e2e": {"value": pairs_per_step * args.steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": int(hq.numel() * 4 + ht.numel() * 4),
                "d2h_bytes_per_step": int(h_idx.numel() * 4 + h_dst.numel() * 4 + h_keep.numel())},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": None,
                     "kernel_ms": k_ms, "kernel_launches": int(kernel_n),
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if pk else "fallback 1.4 PFLOP/s sustained"},
        "clocks": clocks,
    }
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        out = run_reference(args, rank)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    out = run_ours(args, rank, world, device)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cores = cpu_threads()
            hp = importlib.import_module(PKG)
            ns = pick_cpu_sample(args.nt, args.cpu_sample)
            qn, tn = hp.synth.sift_like(ns, args.nt, seed=7, planted=0.1)
            reference_step(qn[:max(16, ns // 8)], tn)
            t0 = time.perf_counter()
            reference_step(qn, tn)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": ns * args.nt / dt, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                   "sample": "%d queries x %d db rows, cv2.BFMatcher(NORM_L2).knnMatch k=2 + ratio loop "
                                             "(features.py:59,:45-55), %.1f s" % (ns, args.nt, dt)}
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
