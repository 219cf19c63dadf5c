#!/usr/bin/env python
"""Headline benchmark of the matching hot path (BASELINE.json): descriptor-pairs/s of brute-force SIFT kNN
(k=2) + Lowe ratio test, 1 M queries x 1 M database rows per GPU (config "synthetic SIFT 128-d float, 1M
query x 1M database kNN k=2 ratio test on 1 B200").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (libhpusm.so through the C ABI)
    python bench.py --impl reference ...                      # the reference's CPU path (cv2.BFMatcher + ratio loop)

One step = one full kNN-2 + ratio pass of the query set over this rank's database shard.  With N > 1 the database
is sharded by rows across ranks, queries are replicated, each rank produces a local top-2, and one NCCL all-gather of
the 16 B/query partial results followed by the lexicographic merge + ratio kernel gives every rank the global answer.
The headline `value` is WEAK scaling (every rank owns `--nt` rows, config.nt_global = N x nt); the same line carries
`strong_scaling` (the fixed nq x nt problem with the database rows divided over the ranks), `nccl_parity` (rank 0
checks the merged answer of a query slice against one single-GPU pass over the gathered database) and `others`: the
remaining BASELINE.json configs (the general-float variant of this one, ORB retrieval, batched RANSAC, pose database), each
with its own value / e2e / roofline / clocks and, at N = 1, cpu_baseline.  Prints ONE JSON line on rank 0.
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


_T0 = time.time()


def log(msg):
    """Progress on stderr (stdout carries exactly one JSON line)."""
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench %6.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sift", choices=["sift", "orb", "ransac", "pose"],
                    help="sift = BASELINE.json's headline config (default); the others are the remaining configs")
    ap.add_argument("--descriptors", default="int", choices=["int", "float"],
                    help="sift: integer-valued SIFT-like rows (the named config) or the general-float variant (x 1.0137)")
    ap.add_argument("--l2-engine", default="auto", choices=["auto", "i8", "bf16", "split", "q16", "simt"],
                    help="sift: force one L2 engine (default: HPUSM_L2_AUTO picks from the data)")
    ap.add_argument("--hamming-engine", default="auto", choices=["auto", "popc", "tc"],
                    help="orb: engine of the segmented Hamming matcher (auto = tcgen05 kind::i8 where its shape limits hold)")
    ap.add_argument("--nq", type=int, default=1 << 20)
    ap.add_argument("--nt", type=int, default=1 << 20, help="database rows PER GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline", default=None, choices=["sift", "orb", "ransac", "pose"],
                    help="internal: print the CPU baseline of one workload (a fresh process, no CUDA) and exit")
    ap.add_argument("--images", type=int, default=10000, help="orb: database images (2048 descriptors each)")
    ap.add_argument("--pairs", type=int, default=100000, help="ransac: image pairs (2000 matches x 1024 hypotheses each)")
    ap.add_argument("--poses", type=int, default=10_000_000, help="pose: database poses")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="sift at N > 1: what `value` reports (the other form is still measured and reported beside it)")
    ap.add_argument("--no-others", action="store_true", help="sift: skip the orb / ransac / pose workloads of the default line")
    ap.add_argument("--min-region-ms", type=float, default=2500.0,
                    help="orb / ransac / pose: steps are raised until the timed region lasts at least this long")
    return ap.parse_args()


def measured_traffic(key, field="bytes"):
    """DRAM bytes per launch of the dominant kernel (or the pipe-utilisation figures, field="ncu") from the committed
    ncu capture of this workload, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key, {}).get(field)
    except Exception:
        return None


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
    """This rank's database shard `t` and the query set `q`, IDENTICAL on every rank (queries are replicated): each rank
    plants noisy copies of rows of its own shard, and the planted rows of all ranks are exchanged and applied in rank
    order, so every rank holds the same queries (round 1 planted per rank and the ranks matched slightly different
    query sets -- found by the on-hardware parity check below)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    t = sift_like_torch(nt, 1000 + rank, device)
    q = sift_like_torch(nq, 7, device)
    k = int(planted * nq) // max(1, world)
    if k:
        g = torch.Generator(device=device).manual_seed(99 + rank)
        qi = torch.randperm(nq, generator=g, device=device)[:k]
        ti = torch.randint(0, nt, (k,), generator=g, device=device)
        noise = torch.round(torch.randn((k, 128), generator=g, device=device) * 4.0)
        rows = torch.clamp(t[ti] + noise, 0, 255)
        if world > 1:
            import torch.distributed as dist
            all_qi = torch.empty((world * k,), dtype=qi.dtype, device=device)
            all_rows = torch.empty((world * k, 128), dtype=rows.dtype, device=device)
            dist.all_gather_into_tensor(all_qi, qi.contiguous())
            dist.all_gather_into_tensor(all_rows, rows.contiguous())
            for r in range(world):
                q[all_qi[r * k:(r + 1) * k]] = all_rows[r * k:(r + 1) * k]
        else:
            q[qi] = rows
    return q, t


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (profiling recipe's clocks line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Samples taken before this call (start-up, warm-up) are dropped from the summary."""
        self.first = len(self.rows)

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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:] or self.rows[-1:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None,
                "power_w_max": max(pw) if pw else None}


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


def load_synth():
    """The synthetic-data generators (numpy only) loaded by path: the reference arm must not import the package, which
    would load libhpusm.so into a process that is supposed to run none of our code."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_hpusm_synth", os.path.join(ROOT, PKG, "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sift_config(args, world):
    """`config` of the sift line: the same keys on both arms."""
    return {"workload": "synthetic SIFT 128-d float, 1M query x 1M database kNN k=2 ratio test on 1 B200",
            "nq": args.nq, "nt_per_gpu": args.nt, "nt_global": args.nt * world, "d": 128, "ratio": 0.8,
            "descriptors": args.descriptors}


def run_reference(args, rank):
    if rank != 0:
        return None
    cores = cpu_threads()
    ns = pick_cpu_sample(args.nt, args.cpu_sample)
    q, t = load_synth().sift_like(ns, args.nt, seed=7, planted=0.1)
    if args.descriptors == "float":
        q, t = q * np.float32(1.0137), t * np.float32(1.0137)
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
        "config": sift_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference",
                         "sample": "%d queries x %d db rows per step, cv2.BFMatcher(NORM_L2).knnMatch k=2 + ratio loop "
                                   "(features.py:59,:45-55)" % (ns, args.nt)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


DTYPE_OF = {
    "tcgen05-i8-exact-int": "u8 operands (exact for SIFT's integers 0..255) / s32 accumulate + f32 re-rank",
    "tcgen05-bf16-exact-int": "bf16 operands (exact for integer-valued SIFT) / f32 accumulate + f32 re-rank",
    "tcgen05-bf16-split": "bf16 hi/lo split operands / f32 accumulate + verified f32 re-rank",
    "tcgen05-i8-q16-split": "15-bit fixed point in two s8 operands / s32 accumulate + verified f32 re-rank",
    "simt-fp32": "f32",
}


# ---- our path ---------------------------------------------------------------------------------------------------------
def sift_pass(args, rank, world, device, hp, nt_rank, row0, label):
    """Times the sharded kNN-2 + ratio pass with `nt_rank` database rows on this rank (global row offset `row0`):
    resident (inputs in HBM) and end to end (host buffers, copies inside the timed region)."""
    import torch.distributed as dist
    eng, rt = hp.engine, hp.retrieval
    nq = args.nq
    q, t = make_workload(nq, nt_rank, rank, device)
    if args.descriptors == "float":  # same rows, no longer integer valued: exercises the split-operand engine
        q, t = (q * 1.0137).contiguous(), (t * 1.0137).contiguous()
    l2_mode = {"auto": eng.L2_AUTO, "i8": eng.L2_TC_I8, "bf16": eng.L2_TC_INT, "split": eng.L2_TC_SPLIT,
               "q16": eng.L2_TC_Q16, "simt": eng.L2_SIMT}[args.l2_engine]
    ws = eng.knn2_l2_workspace(nq, nt_rank, 128, l2_mode, device)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=device)
    dst = torch.empty((nq, 2), dtype=torch.float32, device=device)

    def local_knn(qq, tt):
        return eng.knn2_l2(qq, tt, mode=l2_mode, out=(idx, dst), workspace=ws)

    def step():
        return rt.knn2_sharded(q, t, row0, kind="l2", ratio=0.8, knn_fn=local_knn)

    tm = Timer(args, rank, world, device, hp)
    if label == "strong":   # 1/N of the work per step: raise the step count until the region lasts --min-region-ms
        tm.fit_region(step)
    r = tm.run(step)
    mi, md, keep, count = r["out"]
    survivors = int(count.item())
    mode = {0: "auto", 1: "simt-fp32", 2: "tcgen05-bf16-exact-int", 3: "tcgen05-bf16-split",
            4: "tcgen05-i8-exact-int", 5: "tcgen05-i8-q16-split"}.get(eng.knn2_l2_last_mode(), "?")

    # on-hardware parity of the NCCL path: the merged answer of a query slice == ONE single-GPU pass over the gathered database
    parity = None
    if world > 1 and label == "weak":
        ns = min(4096, nq)
        t_all = rt.allgather(t).reshape(-1, 128)  # every rank takes part in the gather; rank 0 checks
        if rank == 0:
            ri, rd = eng.knn2_l2(q[:ns].contiguous(), t_all, mode=l2_mode)
            ok = bool(torch.equal(ri, mi[:ns]) and torch.equal(rd, md[:ns]))
            parity = {"queries": ns, "db_rows": int(t_all.shape[0]), "identical": ok}
            assert ok, "NCCL all-gather + merge disagrees with the single-GPU pass over the gathered database"
        del t_all
        tm.barrier()

    # end to end through the public API with HOST buffers.  The database shard and this rank's SLICE of the queries are
    # copied host->device (pinned) every step; the query slices are exchanged over NVLink (one all-gather), so each query
    # row crosses PCIe once per node instead of once per rank; the answer is read back every step.
    qa, qb = rt.shard_range(nq, rank, world)
    hq, ht = q[qa:qb].cpu().pin_memory(), t.cpu().pin_memory()
    h_idx = torch.empty((nq, 2), dtype=torch.int32).pin_memory()
    h_dst = torch.empty((nq, 2), dtype=torch.float32).pin_memory()
    h_keep = torch.empty((nq,), dtype=torch.uint8).pin_memory()
    streamed = rt.StreamedL2Matcher(nq, nt_rank, 128, device, mode=l2_mode, query_slices=world > 1)

    def e2e_step():
        gi, gd, kp, _ = rt.knn2_sharded(hq, ht, row0, kind="l2", ratio=0.8, knn_fn=streamed)
        h_idx.copy_(gi, non_blocking=True)
        h_dst.copy_(gd, non_blocking=True)
        h_keep.copy_(kp, non_blocking=True)

    e = tm.run(e2e_step, profile=False)
    assert int(h_keep.sum()) == survivors, "e2e pass disagrees with the resident pass"
    del streamed
    k = float(args.steps) / float(r["steps"])   # the callers' formulas count args.steps steps: scale the regions to that
    return {"ms": r["ms"] * k, "e2e_ms": e["ms"] * k, "kernel_ms": r["kernel_ms"], "kernel_n": r["kernel_n"], "launches": int(r["launches"] * k),
            "steps_timed": int(r["steps"]),
            "clocks": r["clocks"], "survivors": survivors, "mode": mode, "parity": parity, "nt_rank": nt_rank,
            "h2d": int(hq.numel() * 4 + ht.numel() * 4), "d2h": int(h_idx.numel() * 4 + h_dst.numel() * 4 + h_keep.numel()),
            "fallback_rows": eng.knn2_l2_last_fallback_rows(), "answer": (mi.clone(), md.clone())}


def sift_query_sharded_pass(args, rank, world, device, hp, nt_rank, row0, db_sharded_answer):
    """The strong-scaling problem (the fixed nq x nt pairs) with the QUERIES sharded and the database replicated
    (retrieval.knn2_query_sharded): every rank matches its 1/N of the queries against the whole database, the answers are
    all-gathered.  Same data as the row-sharded strong pass (whose shards are gathered into the replica), and the answer
    must be identical to it -- checked here on the hardware."""
    eng, rt = hp.engine, hp.retrieval
    nq = args.nq
    q, t = make_workload(nq, nt_rank, rank, device)
    t_full = rt.allgather(t).reshape(-1, 128).contiguous()
    del t
    qa, qb = rt.shard_range(nq, rank, world)
    q_slice = q[qa:qb].contiguous()
    l2_mode = {"auto": eng.L2_AUTO, "i8": eng.L2_TC_I8, "bf16": eng.L2_TC_INT, "split": eng.L2_TC_SPLIT,
               "q16": eng.L2_TC_Q16, "simt": eng.L2_SIMT}[args.l2_engine]
    ws = eng.knn2_l2_workspace(qb - qa, t_full.shape[0], 128, l2_mode, device)
    idx = torch.empty((qb - qa, 2), dtype=torch.int32, device=device)
    dst = torch.empty((qb - qa, 2), dtype=torch.float32, device=device)

    def local_knn(qq, tt):
        return eng.knn2_l2(qq, tt, mode=l2_mode, out=(idx, dst), workspace=ws)

    def step():
        return rt.knn2_query_sharded(q_slice, t_full, nq, kind="l2", ratio=0.8, knn_fn=local_knn)

    tm = Timer(args, rank, world, device, hp)
    tm.fit_region(step)
    r = tm.run(step)
    mi, md, keep, count = r["out"]
    same = bool(torch.equal(mi, db_sharded_answer[0]) and torch.equal(md, db_sharded_answer[1]))
    assert same, "query-sharded and row-sharded answers differ"
    hq, ht = q_slice.cpu().pin_memory(), t_full.cpu().pin_memory()
    h_idx = torch.empty((nq, 2), dtype=torch.int32).pin_memory()
    h_dst = torch.empty((nq, 2), dtype=torch.float32).pin_memory()
    h_keep = torch.empty((nq,), dtype=torch.uint8).pin_memory()
    streamed = rt.StreamedL2Matcher(qb - qa, t_full.shape[0], 128, device, mode=l2_mode)

    def e2e_step():
        gi, gd, kp, _ = rt.knn2_query_sharded(hq, ht, nq, kind="l2", ratio=0.8, knn_fn=streamed)
        h_idx.copy_(gi, non_blocking=True)
        h_dst.copy_(gd, non_blocking=True)
        h_keep.copy_(kp, non_blocking=True)

    e = tm.run(e2e_step, profile=False)
    assert int(h_keep.sum()) == int(count.item()), "e2e pass disagrees with the resident pass"
    del streamed
    k = float(args.steps) / float(r["steps"])
    return {"ms": r["ms"] * k, "e2e_ms": e["ms"] * k, "kernel_ms": r["kernel_ms"], "kernel_n": r["kernel_n"], "identical": same,
            "steps_timed": int(r["steps"]), "nt_rank": int(t_full.shape[0]), "h2d": int(hq.numel() * 4 + ht.numel() * 4)}


_I8_PEAK = {}


def i8_sustained_peak(hp, device, seconds=2.0):
    """Sustained issue rate of tcgen05.mma kind::i8 on this GPU (T op/s): the roofline denominator of the integer tensor-core
    kernels, measured in this run -- back-to-back MMAs on every SM for `seconds`, clocks sampled beside it
    (MEASURED_PEAKS.json has no INT8 line).  Measured once per process."""
    if "v" not in _I8_PEAK:
        _I8_PEAK["v"] = _i8_sustained_peak(hp, device, seconds)
    return _I8_PEAK["v"]


def _i8_sustained_peak(hp, device, seconds):
    eng = hp.engine
    rounds = 4096  # ~10 ms per launch
    eng.microbench_mma("i8", 64, device)
    torch.cuda.synchronize(device)
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, ops, t0 = 0, 0.0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        ops = eng.microbench_mma("i8", rounds, device)
        n += 1
        if n % 8 == 0:
            torch.cuda.synchronize(device)
    e1.record()
    torch.cuda.synchronize(device)
    clocks = sampler.stop()
    return ops * n / (e0.elapsed_time(e1) * 1e-3) / 1e12, clocks


def run_sift_float(args, rank, world, device, hp, i8_peak=None):
    """SURVEY 8d's "general float" variant of C2 (the same rows x 1.0137: no longer integers) as one of `others`:
    HPUSM_L2_AUTO takes the fixed-point integer engine (l2_tc_q16.cu).  Weak form at N > 1 (1 M database rows per GPU)."""
    import copy
    a = copy.copy(args)
    a.descriptors, a.l2_engine, a.steps = "float", "auto", max(3, min(args.steps, 10))
    r = sift_pass(a, rank, world, device, hp, a.nt, rank * a.nt, "weak-float")
    r.pop("answer", None)
    if rank != 0:
        return None
    pairs = float(a.nq) * a.nt * world
    k_ms = r["kernel_ms"] / max(r["kernel_n"], 1)
    achieved = FLOP_PER_PAIR * float(a.nq) * a.nt * a.steps / max(r["kernel_n"], 1) / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    peak = i8_peak or 0.0
    key = "sift:%s:%dx%d" % (r["mode"], a.nq, a.nt)
    return {"metric": "descriptor-pairs/sec (kNN+ratio)", "value": pairs * a.steps / (r["ms"] * 1e-3), "unit": "pairs/s", "n_gpus": world,
            "steps": a.steps, "ms_per_step": r["ms"] / a.steps, "scaling": "weak", "dtype": DTYPE_OF.get(r["mode"], "f32"), "data": "synthetic",
            "config": dict(sift_config(a, world), workload="C2 general-float variant: the same synthetic rows x 1.0137 (not integers), 1M x 1M kNN k=2 ratio test"),
            "detail": {"engine": r["mode"], "fallback_rows": r["fallback_rows"], "ratio_survivors": r["survivors"]},
            "e2e": {"value": pairs * a.steps / (r["e2e_ms"] * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": int(r["launches"]),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "frac_issued": achieved / peak * 832.0 / FLOP_PER_PAIR if peak and r["mode"] == "tcgen05-i8-q16-split" else None,
                         "traffic": measured_traffic(key), "kernel_ms": k_ms, "kernel_launches": int(r["kernel_n"]),
                         "peak_source": "the sustained kind::i8 issue rate measured for the headline line of this run; frac counts the 256 "
                                        "algorithmic op per pair, frac_issued the 832 the 13 MMAs per tile issue",
                         "ncu": measured_traffic(key, "ncu")},
            "clocks": r["clocks"]}


def run_sift(args, rank, world, device, hp):
    log("sift: weak-scaling pass (resident + end to end)")
    weak = sift_pass(args, rank, world, device, hp, args.nt, rank * args.nt, "weak")
    strong = weak
    if world > 1:
        a, b = hp.retrieval.shard_range(args.nt, rank, world, align=128)
        log("sift: strong-scaling pass")
        strong = sift_pass(args, rank, world, device, hp, b - a, a, "strong")
        log("sift: strong-scaling pass, queries sharded / database replicated")
        strong_q = sift_query_sharded_pass(args, rank, world, device, hp, b - a, a, strong["answer"])
    weak.pop("answer", None)
    strong.pop("answer", None)
    if rank != 0:
        return None
    nq, nt = args.nq, args.nt
    head, other = (weak, strong) if args.scaling == "weak" else (strong, weak)
    pairs_weak, pairs_strong = float(nq) * nt * world, float(nq) * nt
    pairs_head = pairs_weak if args.scaling == "weak" else pairs_strong
    value = pairs_head * args.steps / (head["ms"] * 1e-3)
    pk = peaks()
    bf16_tf = (pk or {}).get("bf16_tflops_sustained", 1400.0)
    k_ms = head["kernel_ms"] / max(head["kernel_n"], 1)
    flops_per_launch = FLOP_PER_PAIR * float(nq) * head["nt_rank"] * head.get("steps_timed", args.steps) / max(head["kernel_n"], 1)
    achieved = flops_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    mode = head["mode"]
    peak_tf, peak_clocks = bf16_tf, None
    peak_source = "MEASURED_PEAKS.json bf16_tflops_sustained" if pk else "fallback 1.4 PFLOP/s sustained"
    if mode in ("tcgen05-i8-exact-int", "tcgen05-i8-q16-split"):
        # MEASURED_PEAKS.json has no INT8 line: the denominator is this GPU's own sustained kind::i8 issue rate, measured now
        log("sift: sustained kind::i8 issue rate (roofline denominator)")
        peak_tf, peak_clocks = i8_sustained_peak(hp, device)
        peak_source = ("hpusm_microbench_mma(kind::i8) sustained for 2 s in this run: every SM issuing back-to-back 128x256x32 "
                       "u8 MMAs and nothing else (clocks under that load in roofline.peak_clocks)")
    out = {
        "metric": "descriptor-pairs/sec (kNN+ratio)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms"] / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": DTYPE_OF.get(mode, "f32"), "data": "synthetic",
        "config": sift_config(args, world),
        "detail": {"engine": mode, "fallback_rows": head["fallback_rows"], "sharding": "database rows",
                   "l2_flush": "inputs (2 x %.0f MB) exceed the 126 MB L2" % (nq * 512 / 1e6), "ratio_survivors": head["survivors"]},
        "e2e": {"value": pairs_head * args.steps / (head["e2e_ms"] * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"],
                "note": "per rank: its database shard + its 1/N slice of the queries over PCIe, query slices all-gathered over NVLink"},
        "gpu_launches": int(head["launches"]),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None, "traffic": measured_traffic("sift:%s:%dx%d" % (mode, nq, nt)),
                     "kernel_ms": k_ms, "kernel_launches": int(head["kernel_n"]),
                     "frac_of_bf16_peak": achieved / bf16_tf if bf16_tf else None, "peak_source": peak_source,
                     "peak_clocks": peak_clocks, "ncu": measured_traffic("sift:%s:%dx%d" % (mode, nq, nt), "ncu")},
        "clocks": head["clocks"],
    }
    if mode == "tcgen05-i8-q16-split" and out["roofline"]["frac"]:
        # general floats cost 13 kind::i8 MMAs of K = 32 per tile where the integer engine needs 5: 832 issued op per pair
        # for 256 algorithmic ones.  `frac` stays algorithmic; this is the share of the measured issue rate the kernel uses.
        out["roofline"]["frac_issued"] = out["roofline"]["frac"] * 832.0 / FLOP_PER_PAIR
    for name, res, pairs in (("weak_scaling", weak, pairs_weak), ("strong_scaling", strong, pairs_strong)):
        out[name] = {"value": pairs * args.steps / (res["ms"] * 1e-3), "unit": "pairs/s", "ms_per_step": res["ms"] / args.steps,
                     "e2e": pairs * args.steps / (res["e2e_ms"] * 1e-3), "nq": nq, "nt_global": int(pairs / nq),
                     "nt_per_gpu": res["nt_rank"], "kernel_ms": res["kernel_ms"] / max(res["kernel_n"], 1),
                     "note": "efficiency at N GPUs = value(N) / (N x value(1)) for weak, value(N) / (N x value(1)) for strong "
                             "(strong: the same nq x nt pairs in 1/N of the time)"}
    if world > 1:
        out["strong_scaling_query_sharded"] = {
            "value": pairs_strong * args.steps / (strong_q["ms"] * 1e-3), "unit": "pairs/s", "ms_per_step": strong_q["ms"] / args.steps,
            "e2e": pairs_strong * args.steps / (strong_q["e2e_ms"] * 1e-3), "nq_per_gpu": -(-nq // world), "nt_per_gpu": strong_q["nt_rank"],
            "kernel_ms": strong_q["kernel_ms"] / max(strong_q["kernel_n"], 1), "h2d_bytes_per_step": strong_q["h2d"],
            "steps_timed": strong_q["steps_timed"],
            "identical_to_row_sharded": strong_q["identical"],
            "note": "same nq x nt pairs, retrieval.knn2_query_sharded: 1/N of the queries per rank against the replicated database, "
                    "one all-gather of the answers; the form for databases that fit every GPU"}
    if weak["parity"] is not None:
        out["nccl_parity"] = weak["parity"]
    return out


# ---- the remaining configs of BASELINE.json (run with --workload; same JSON contract) -----------------------------------
class Timer:
    """W warm-up steps, then K timed steps between CUDA events with a barrier + synchronize on both sides; MAX over ranks."""

    def __init__(self, args, rank, world, device, hp):
        self.args, self.rank, self.world, self.device, self.hp = args, rank, world, device, hp
        self.steps = args.steps

    def fit_region(self, step):
        """Raises the step count until the timed region lasts --min-region-ms (short steps: >= 20 clock samples)."""
        step()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        self.barrier()
        ms = max(self.max_over_ranks(e0.elapsed_time(e1)), 1e-3)
        self.steps = int(max(self.args.steps, min(100000, -(-self.args.min_region_ms // ms))))
        return self.steps

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.device)

    def max_over_ranks(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def run(self, step, profile=True):
        eng = self.hp.engine
        sampler = ClockSampler(self.device.index or 0)
        if self.rank == 0 and profile:
            sampler.start()
        for _ in range(self.args.warmup):
            step()
        self.barrier()
        sampler.mark()
        if profile:
            eng.profile_enable(True)
            eng.profile_collect()
        l0 = self.hp.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(self.steps):
            out = step()
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        res = {"ms": ms, "out": out, "launches": self.hp.launch_count() - l0, "steps": self.steps}
        if profile:
            res["kernel_ms"], res["kernel_n"] = eng.profile_collect()
            eng.profile_enable(False)
            res["clocks"] = sampler.stop() if self.rank == 0 else None
        return res


def base_line(args, world, metric, unit, value, ms, dtype, scaling, config, steps=None):
    steps = steps or args.steps
    return {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": config}


def popc_peak(hp, device, op="popc"):
    """Measured POPC.32 / LOP3 issue rate of this GPU (thread-instructions / s): the Hamming roofline denominators."""
    eng = hp.engine
    blocks, threads, iters = 148 * 16, 256, 1 << 16
    eng.microbench_popc(blocks, threads, iters, device, op)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.microbench_popc(blocks, threads, iters, device, op)
    e1.record()
    torch.cuda.synchronize(device)
    return blocks * threads * float(iters) / (e0.elapsed_time(e1) * 1e-3)


def run_orb(args, rank, world, device, hp):
    """C3: one 2048-descriptor query image against a 10k-image ORB database sharded by images (strong scaling)."""
    eng, rt = hp.engine, hp.retrieval
    per = 2048
    a, b = rt.shard_range(args.images, rank, world)
    g = torch.Generator(device=device).manual_seed(1)
    query = torch.randint(0, 256, (per, 32), generator=g, device=device, dtype=torch.uint8)
    g2 = torch.Generator(device=device).manual_seed(100 + rank)
    db = torch.randint(0, 256, ((b - a) * per, 32), generator=g2, device=device, dtype=torch.uint8)
    planted = list(range(a, b, 20))  # 5 % of the images hold perturbed copies of half the query descriptors
    for im in planted:
        rows = torch.arange(0, per, 2, device=device)
        flips = (torch.rand((rows.numel(), 32, 8), generator=g2, device=device) < 12.0 / 256).to(torch.uint8)
        mask = (flips * (1 << torch.arange(8, device=device, dtype=torch.uint8))).sum(-1).to(torch.uint8)
        db[(im - a) * per + rows] = query[rows] ^ mask
    offs = (torch.arange(0, b - a + 1, device=device, dtype=torch.int64) * per)
    tm = Timer(args, rank, world, device, hp)

    heng = args.hamming_engine

    def step():
        return rt.image_scores_sharded(query, db, offs, 0.8, hamming_engine=heng)

    tm.fit_region(step)
    r = tm.run(step)
    scores = r["out"]
    hq, hdb = query.cpu().pin_memory(), db.cpu().pin_memory()
    hs = torch.empty((args.images,), dtype=torch.int32).pin_memory()
    # the public host-array entry point: runs of whole images uploaded on a copy stream while earlier runs are scored
    streamed = rt.StreamedImageScorer(per, 32, offs, device, hamming_engine=heng)

    def e2e_step():
        hs.copy_(rt.image_scores_sharded(hq, hdb, offs, 0.8, score_fn=streamed), non_blocking=True)

    e = tm.run(e2e_step, profile=False)
    if rank != 0:
        return None
    pairs = float(per) * args.images * per
    top = torch.topk(scores, min(len(planted) * world, 16)).indices.tolist()
    k_ms = r["kernel_ms"] / max(r["kernel_n"], 1)
    pairs_per_launch = float(per) * (b - a) * per * tm.steps / max(r["kernel_n"], 1)
    tensor = heng != "popc"   # at this shape AUTO resolves to the tensor-core engine
    out = base_line(args, world, "descriptor-pairs/sec (kNN+ratio)", "pairs/s", pairs * tm.steps / (r["ms"] * 1e-3), r["ms"],
                    "s8 x s8 -> s32 (tcgen05 kind::i8)" if tensor else "u32 popc",
                    "strong", {"workload": "synthetic ORB 256-bit Hamming, 10k-image database retrieval sharded across 2/4/8 B200",
                               "images": args.images, "descriptors_per_image": per, "query_descriptors": per, "ratio": 0.8,
                               "sharding": "database images", "hamming_engine": "tc" if tensor else "popc",
                               "l2_flush": "database shard (%.0f MB) exceeds the 126 MB L2" % (db.numel() / 1e6),
                               "top_images": top[:8]}, steps=tm.steps)
    out["e2e"] = {"value": pairs * tm.steps / (e["ms"] * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(hq.numel() + hdb.numel()),
                  "d2h_bytes_per_step": int(hs.numel() * 4)}
    out["gpu_launches"] = int(r["launches"])
    if tensor:
        # ALGORITHMIC operations: 256 bit comparisons per pair = 256 multiply-accumulates = 512 operations; the engine issues
        # 288 (32 augmentation columns carry the row index and the padding lift), which the fraction does not credit
        i8_peak, i8_clocks = i8_sustained_peak(hp, device)
        ach = 512.0 * pairs_per_launch / (k_ms * 1e-3) / 1e12
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": i8_peak, "unit": "T op/s (s8 x s8 -> s32)", "frac": ach / i8_peak,
                           "traffic": measured_traffic("orb_tc:%dx%d" % (args.images, world)), "kernel_ms": k_ms,
                           "kernel_launches": int(r["kernel_n"]), "step_ms": r["ms"] / tm.steps,
                           "peak_source": "hpusm_microbench_mma(i8) sustained for 2 s in this run; the step also holds the operand "
                                          "expansion (32 -> 288 bytes per database row) and the plan kernels, outside kernel_ms",
                           "peak_clocks": i8_clocks, "ops_per_pair": 512, "issued_ops_per_pair": 576, "frac_of_issued": ach / i8_peak * 576.0 / 512.0,
                           "hbm_GBps_algorithmic": (db.numel() * 9 + query.numel()) / (k_ms * 1e-3) / 1e9}
        if world == 1:
            # the __popc engine (north_star's integer-pipe form, and the cross-check of the tensor engine) on the same inputs
            tm2 = Timer(args, rank, world, device, hp)
            tm2.steps = max(3, min(args.steps, 10))
            r2 = tm2.run(lambda: rt.image_scores_sharded(query, db, offs, 0.8, hamming_engine="popc"))
            assert torch.equal(r2["out"], scores), "the two Hamming engines disagree"
            ppeak, lpeak = popc_peak(hp, device, "popc"), popc_peak(hp, device, "lop3")
            k2 = r2["kernel_ms"] / max(r2["kernel_n"], 1) * 1e-3
            out["popc_engine"] = {"value": pairs * tm2.steps / (r2["ms"] * 1e-3), "unit": "pairs/s", "ms_per_step": r2["ms"] / tm2.steps,
                                  "identical_scores": True,
                                  "roofline": {"bound": "integer pipes (16 LOP3 + 4 POPC.32 per pair)",
                                               "frac": max(16.0 * pairs / lpeak, 4.0 * pairs / ppeak) / k2, "kernel_ms": k2 * 1e3,
                                               "traffic": measured_traffic("orb:%dx%d" % (args.images, world))}}
    else:
        ppeak, lpeak = popc_peak(hp, device, "popc"), popc_peak(hp, device, "lop3")
        # per 256-bit pair the kernel issues 16 LOP3 (8 XOR + 4 carry-save adders) on the ALU pipe and 4 POPC on the XU pipe;
        # the two pipes run side by side, so the bound is the slower of the two
        t_bound = max(16.0 * pairs_per_launch / lpeak, 4.0 * pairs_per_launch / ppeak)
        ach = 16.0 * pairs_per_launch / (k_ms * 1e-3)
        out["roofline"] = {"bound": "integer pipes (16 LOP3 + 4 POPC.32 per pair)", "achieved": ach / 1e12, "peak": lpeak / 1e12, "unit": "Tlop3/s",
                           "frac": t_bound / (k_ms * 1e-3), "traffic": measured_traffic("orb:%dx%d" % (args.images, world)), "kernel_ms": k_ms, "kernel_launches": int(r["kernel_n"]),
                           "peak_source": "hpusm_microbench_lop3 / hpusm_microbench_popc measured in this run",
                           "popc_peak_Tpopc": ppeak / 1e12, "naive_popc_bound_ms": 8.0 * pairs_per_launch / ppeak * 1e3,
                           "hbm_GBps_algorithmic": (db.numel() + query.numel()) / (k_ms * 1e-3) / 1e9}
    out["clocks"] = r["clocks"]
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_subprocess("orb")
    return out


def run_ransac(args, rank, world, device, hp):
    """C4: batched RANSAC homography scoring, pairs sharded over ranks (independent units, no collective): hyps/s."""
    eng, rt = hp.engine, hp.retrieval
    K, nhyp = 2000, 1024
    a, b = rt.shard_range(args.pairs, rank, world)
    n = b - a
    g = torch.Generator(device=device).manual_seed(2 + rank)
    th = (torch.rand(n, generator=g, device=device) - 0.5) * (3.14159265 / 3)
    sc = 0.7 + 0.7 * torch.rand(n, generator=g, device=device)
    H = torch.zeros((n, 3, 3), device=device)
    H[:, 0, 0] = sc * torch.cos(th); H[:, 0, 1] = -sc * torch.sin(th); H[:, 1, 0] = sc * torch.sin(th); H[:, 1, 1] = sc * torch.cos(th)
    H[:, :2, 2] = (torch.rand((n, 2), generator=g, device=device) - 0.5) * 80
    H[:, 2, :2] = (torch.rand((n, 2), generator=g, device=device) - 0.5) * 6e-4
    H[:, 2, 2] = 1
    src = torch.rand((n, K, 2), generator=g, device=device) * torch.tensor([640.0, 480.0], device=device)
    p = torch.einsum("pij,pnj->pni", H, torch.cat([src, torch.ones((n, K, 1), device=device)], 2))
    dst = p[..., :2] / p[..., 2:]
    inl = torch.rand((n, K), generator=g, device=device) < 0.3
    dst = torch.where(inl[..., None], dst + torch.randn((n, K, 2), generator=g, device=device),
                      torch.rand((n, K, 2), generator=g, device=device) * torch.tensor([640.0, 480.0], device=device))
    src, dst = src.reshape(-1, 2).contiguous(), dst.reshape(-1, 2).contiguous()
    del p, inl, H
    offs = torch.arange(0, n + 1, device=device, dtype=torch.int64) * K
    counts = torch.empty((n, nhyp), dtype=torch.int32, device=device)
    tm = Timer(args, rank, world, device, hp)

    def step():
        return eng.ransac_score_batch(src, dst, offs, 5.0, nhyp, seed=2, out=counts, pair_index0=a)  # a = this shard's first pair

    tm.fit_region(step)
    r = tm.run(step)
    scored_t = (counts >= 0).sum().to(torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(scored_t)
    scored_all = float(scored_t.item())        # hypotheses whose model was scored, over all ranks (every one of them, by design)
    scored = int((counts >= 0).sum().item())
    best = counts.max(dim=1).values.float().mean().item()
    # end to end = the full findHomography replacement (score + early-exit replay + refinement + mask) on a host batch
    ne = min(n, 4096)
    hs, hd = src[:ne * K].cpu().pin_memory(), dst[:ne * K].cpu().pin_memory()
    hH, hm = torch.empty((ne, 3, 3), dtype=torch.float64).pin_memory(), torch.empty((ne * K,), dtype=torch.uint8).pin_memory()
    # the public host-array entry point: runs of whole pairs uploaded on a copy stream while earlier runs are solved
    streamed = hp.retrieval.StreamedHomographyBatch(offs[:ne + 1], device, 5.0, nhyp, 0.0, seed=2)

    def e2e_step():
        Hh, mask, ninl = streamed(hs, hd)
        hH.copy_(Hh, non_blocking=True)
        hm.copy_(mask, non_blocking=True)

    e = tm.run(e2e_step, profile=False)
    # the same full replacement with the matches already resident (no copies): score + replay + refinement + mask
    nf = min(n, 16384)
    Hf, mf, nif = [None], [None], [None]

    def full_step():
        Hf[0], mf[0], nif[0], _ = eng.ransac_homography_batch(src[:nf * K], dst[:nf * K], offs[:nf + 1], 5.0, nhyp, 0.0, seed=2,
                                                              pair_index0=a)

    f = tm.run(full_step, profile=False)
    if rank != 0:
        return None
    hyps = scored_all   # `value` counts SCORED hypotheses only (round 1 counted the 79 % rejected as degenerate too)
    k_ms = r["kernel_ms"] / max(r["kernel_n"], 1)
    evals = float(scored) * K  # reprojections evaluated on this rank per step
    flops = evals * 22.0
    out = base_line(args, world, "RANSAC hyps/sec", "hyps/s", hyps * tm.steps / (r["ms"] * 1e-3), r["ms"], "f32 score / f64 solve", "strong",
                    {"workload": "batched RANSAC homography: 100k image pairs x 2k putative matches x 1024 hypotheses", "pairs": args.pairs,
                     "matches_per_pair": K, "hypotheses": nhyp, "thresh": 5.0, "inlier_ratio": 0.3, "sharding": "image pairs",
                     "sampler": "counter", "scored_fraction": scored / float(n * nhyp), "mean_best_inliers": best,
                     "l2_flush": "match coordinates (%.1f GB) exceed the 126 MB L2" % (src.numel() * 8 / 1e9)}, steps=tm.steps)
    out["e2e"] = {"value": ne * world * nhyp * tm.steps / (e["ms"] * 1e-3), "unit": "hyps/s", "h2d_bytes_per_step": int(hs.numel() * 8),
                  "d2h_bytes_per_step": int(hH.numel() * 8 + hm.numel()), "note": "full findHomography replacement incl. refinement, %d pairs per step" % ne}
    out["find_homography_resident"] = {"value": nf * world * nhyp * tm.steps / (f["ms"] * 1e-3), "unit": "hyps/s", "pairs_per_step": nf,
                                       "note": "score + early-exit replay + DLT/LM refinement + final mask, matches resident in HBM"}
    out["gpu_launches"] = int(r["launches"])
    # measured FP32 peak of this GPU: scalar FFMA (2 flop per lane-instruction) and packed FFMA2 (4 flop), whichever is higher
    f1, f2 = 2.0 * popc_peak(hp, device, "ffma"), 4.0 * popc_peak(hp, device, "ffma2")
    fpeak = max(f1, f2)
    out["roofline"] = {"bound": "fp32 pipe", "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": fpeak / 1e12, "unit": "TFLOP/s",
                       "frac": flops / (k_ms * 1e-3) / fpeak, "traffic": measured_traffic("ransac:%dx%d" % (args.pairs, world)), "kernel_ms": k_ms,
                       "kernel_launches": int(r["kernel_n"]), "peak_source": "hpusm_microbench_ffma measured in this run: scalar FFMA %.1f TFLOP/s, packed FFMA2 %.1f TFLOP/s (MEASURED_PEAKS.json has no FP32 line)" % (f1 / 1e12, f2 / 1e12),
                       "hbm_GBps_algorithmic": src.numel() * 8 / (k_ms * 1e-3) / 1e9, "reprojections_per_s": evals / (k_ms * 1e-3)}
    out["clocks"] = r["clocks"]
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_subprocess("ransac")
    return out


def run_pose(args, rank, world, device, hp):
    """C5: one query pose against a 10 M-pose database sharded by rows; per-rank top-k + one small all-gather."""
    eng, rt = hp.engine, hp.retrieval
    a, b = rt.shard_range(args.poses, rank, world)
    n = b - a
    g = torch.Generator(device=device).manual_seed(3 + rank)
    base = torch.tensor(hp.synth.CANONICAL_POSE, dtype=torch.float32, device=device)
    th = (torch.rand(n, generator=g, device=device) - 0.5) * (50 * 3.14159265 / 180)
    sc = 0.5 + 1.5 * torch.rand(n, generator=g, device=device)
    R = torch.stack([torch.stack([torch.cos(th), -torch.sin(th)], 1), torch.stack([torch.sin(th), torch.cos(th)], 1)], 1) * sc[:, None, None]
    c = base.mean(0)
    db = torch.einsum("nij,kj->nki", R, base - c) + c + (torch.rand((n, 1, 2), generator=g, device=device) - 0.5) * 100 + 400
    sigma = torch.tensor([0.0, 2.0, 5.0, 15.0], device=device)[torch.randint(0, 4, (n,), generator=g, device=device)]
    db = torch.clamp(db + torch.randn((n, 18, 2), generator=g, device=device) * sigma[:, None, None], min=1.0)
    drop = (torch.rand((n, 18), generator=g, device=device) < 0.15) & (torch.rand((n, 1), generator=g, device=device) < 0.1)
    db = (db * (~drop)[..., None]).contiguous()
    del R, drop, sigma
    query = base.clone()
    match = torch.empty((n,), dtype=torch.uint8, device=device)
    score = torch.empty((n,), dtype=torch.float32, device=device)
    tm = Timer(args, rank, world, device, hp)

    def match_fn(qq, dd):
        return eng.pose_match_batch(qq, dd, out=(match, score))

    def step():
        return rt.pose_topk_sharded(query, db, a, k=16, match_fn=match_fn)

    tm.fit_region(step)
    r = tm.run(step)
    ts, ti, cnt = r["out"]
    ne = min(n, 2_000_000)
    hdb = db[:ne].cpu().pin_memory()
    ddb = torch.empty_like(db[:ne])
    hts, hti = torch.empty((16,), dtype=torch.float32).pin_memory(), torch.empty((16,), dtype=torch.int64).pin_memory()

    def e2e_step():
        ddb.copy_(hdb, non_blocking=True)
        s_, i_, _ = rt.pose_topk_sharded(query, ddb, a, k=16)
        hts.copy_(s_, non_blocking=True)
        hti.copy_(i_, non_blocking=True)

    e = tm.run(e2e_step, profile=False)
    if rank != 0:
        return None
    k_ms = r["kernel_ms"] / max(r["kernel_n"], 1)
    pk = peaks()
    peak = (pk or {}).get("hbm_gbs", 6650.0)
    bytes_per_launch = n * (144.0 + 5.0) * tm.steps / max(r["kernel_n"], 1)
    ach = bytes_per_launch / (k_ms * 1e-3) / 1e9
    out = base_line(args, world, "pose pairs/sec (match_single)", "poses/s", float(args.poses) * tm.steps / (r["ms"] * 1e-3), r["ms"], "f64", "strong",
                    {"workload": "OpenPose 18-joint Procrustes/affine pose matching: 1 query vs 10M synthetic pose database, multi-GPU",
                     "poses": args.poses, "matches": int(cnt.item()), "best_score": float(ts[0].item()), "sharding": "database poses",
                     "l2_flush": "database shard (%.2f GB) exceeds the 126 MB L2" % (db.numel() * 4 / 1e9)}, steps=tm.steps)
    out["e2e"] = {"value": ne * world * tm.steps / (e["ms"] * 1e-3), "unit": "poses/s", "h2d_bytes_per_step": int(hdb.numel() * 4),
                  "d2h_bytes_per_step": 16 * 12, "note": "%d poses per step from pinned host memory" % ne}
    out["gpu_launches"] = int(r["launches"])
    out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": measured_traffic("pose:%dx%d" % (args.poses, world)), "kernel_ms": k_ms,
                       "kernel_launches": int(r["kernel_n"]), "peak_source": "MEASURED_PEAKS.json hbm_gbs" if pk else "fallback 6650 GB/s"}
    out["clocks"] = r["clocks"]
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_subprocess("pose")
    return out


# ---- CPU baselines: the reference's own CPU path, each in a FRESH process -------------------------------------------
# (A process that has initialised CUDA, started OpenCV's thread pool and an nvidia-smi reader thread must not fork worker
# pools: the children inherit locked mutexes and hang.  The baselines therefore run in a clean interpreter, with a
# hard timeout, and only report a number.)
def cpu_baseline_subprocess(workload, extra=(), timeout=240):
    log("cpu baseline of %s in a fresh process" % workload)
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-baseline", workload] + list(extra)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
        if res.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": "cpu baseline exited %d: %s" % (res.returncode, res.stderr[-300:])}
    except subprocess.TimeoutExpired:
        return {"error": "cpu baseline timed out after %d s" % timeout}


def _cv2_ransac_pairs(job):
    import cv2
    src, dst = job
    cv2.setNumThreads(1)
    for s, d in zip(src, dst):
        cv2.findHomography(s.reshape(-1, 1, 2), d.reshape(-1, 1, 2), cv2.RANSAC, 5.0, maxIters=1024, confidence=1 - 1e-9)
    return len(src)


def _pose_chunk(job):
    from oracle import pose_oracle
    q, db = job
    return pose_oracle.match_single_batch(q, db)[0].sum()


def run_cpu_baseline(args):
    """`bench.py --cpu-baseline W`: times the reference's CPU implementation of workload W on a bounded sample, all host cores."""
    synth = load_synth()
    cores = os.cpu_count() or 1
    w = args.cpu_baseline
    if w == "sift":
        threads = cpu_threads()
        ns = pick_cpu_sample(args.nt, args.cpu_sample)
        qn, tn = synth.sift_like(ns, args.nt, seed=7, planted=0.1)
        if args.descriptors == "float":
            qn, tn = qn * np.float32(1.0137), tn * np.float32(1.0137)
        reference_step(qn[:max(16, ns // 8)], tn)
        t0 = time.perf_counter()
        reference_step(qn, tn)
        dt = time.perf_counter() - t0
        return {"value": ns * args.nt / dt, "unit": "pairs/s", "cores": threads, "kind": "reference",
                "sample": "%d queries x %d db rows, cv2.BFMatcher(NORM_L2).knnMatch k=2 + ratio loop "
                          "(features.py:59,:45-55), %.1f s" % (ns, args.nt, dt)}
    if w == "orb":
        import cv2
        threads = cpu_threads()
        per, nim = 2048, 1024
        rng = np.random.default_rng(1)
        qn = rng.integers(0, 256, (per, 32), dtype=np.uint8)
        dn = rng.integers(0, 256, (nim * per, 32), dtype=np.uint8)
        m = cv2.BFMatcher(cv2.NORM_HAMMING)
        m.knnMatch(qn, trainDescriptors=dn[:per], k=2)
        t0 = time.perf_counter()
        for i in range(nim):
            raw = m.knnMatch(qn, trainDescriptors=dn[i * per:(i + 1) * per], k=2)
            sum(1 for x in raw if len(x) == 2 and x[0].distance < x[1].distance * 0.8)
        dt = time.perf_counter() - t0
        return {"value": per * per * nim / dt, "unit": "pairs/s", "cores": threads, "kind": "reference",
                "sample": "%d database images x 2048 descriptors, cv2.BFMatcher(NORM_HAMMING).knnMatch k=2 + ratio loop per image, %.1f s" % (nim, dt)}
    import multiprocessing as mpx
    if w == "ransac":
        K, nhyp, npair = 2000, 1024, 96 * cores
        src, dst, _ = synth.ransac_batch(npair, K, 0.3, seed=2)
        sn, dn = src.reshape(npair, K, 2), dst.reshape(npair, K, 2)
        jobs = [(sn[i::cores], dn[i::cores]) for i in range(cores)]
        with mpx.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            pool.map(_cv2_ransac_pairs, jobs)
            dt = time.perf_counter() - t0
        return {"value": npair * nhyp / dt, "unit": "hyps/s", "cores": cores, "kind": "reference",
                "sample": "%d pairs x 2000 matches, cv2.findHomography(RANSAC, 5.0, maxIters=1024, confidence=1-1e-9: every "
                          "iteration scores a model) in a %d-process pool, %.1f s" % (npair, cores, dt)}
    if w == "pose":
        ns = 15000 * cores
        qn = synth.CANONICAL_POSE.astype(np.float64)
        dn = synth.pose_database(synth.CANONICAL_POSE, ns, seed=3).astype(np.float64)
        jobs = [(qn, dn[i::cores]) for i in range(cores)]
        with mpx.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            pool.map(_pose_chunk, jobs)
            dt = time.perf_counter() - t0
        return {"value": ns / dt, "unit": "poses/s", "cores": cores, "kind": "port",
                "sample": "%d poses, oracle/pose_oracle.match_single (NumPy restatement of single_person.py:80-189) in a "
                          "%d-process pool, %.1f s" % (ns, cores, dt)}
    raise SystemExit("unknown --cpu-baseline workload %r" % w)


def main():
    args = parse_args()
    # a wedged run must fail loudly, not sit on the GPU box until somebody's limit kills it
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("HPUSM_BENCH_WATCHDOG_S", "1500")), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_baseline:
        print(json.dumps(run_cpu_baseline(args)), flush=True)
        return
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
    hp = importlib.import_module(PKG)
    runners = {"orb": run_orb, "ransac": run_ransac, "pose": run_pose}
    if args.workload != "sift":
        out = runners[args.workload](args, rank, world, device, hp)
    else:
        out = run_sift(args, rank, world, device, hp)
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            log("sift: cpu baseline")
            out["cpu_baseline"] = cpu_baseline_subprocess("sift", ["--nt", str(args.nt), "--cpu-sample", str(args.cpu_sample),
                                                                    "--descriptors", args.descriptors])
        if not args.no_others:
            others = {}
            i8_peak = (out or {}).get("roofline", {}).get("peak") if rank == 0 and (out or {}).get("detail", {}).get("engine") == "tcgen05-i8-exact-int" else None
            runners["sift_float"] = lambda *a: run_sift_float(*a, i8_peak=i8_peak)
            for name in ("sift_float", "orb", "ransac", "pose"):
                if name == "sift_float" and args.descriptors == "float":
                    continue
                torch.cuda.empty_cache()
                log("others: %s" % name)
                try:
                    res = runners[name](args, rank, world, device, hp)
                except Exception as exc:  # one workload failing must not cost the headline line
                    res = {"error": "%s: %s" % (type(exc).__name__, exc)}
                if rank == 0:
                    others[name] = res
            if rank == 0:
                out["others"] = others
    log("done")
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
