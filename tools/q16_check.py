"""Development check of the fixed-point integer engine for general floats (HPUSM_L2_TC_Q16): parity against the exact
FP32 engine on small shapes, fallback counts, and timings against the bf16 split engine.  Run on the GPU box:
    timeout 600 python tools/q16_check.py [--big]
"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
E = hp.engine
dev = torch.device("cuda:0")


def d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def check(name, q, t):
    dq, dt = d(q), d(t)
    ia, da = hp.knn2_l2(dq, dt, mode=E.L2_TC_Q16)
    fb = E.knn2_l2_last_fallback_rows()
    ib, db = hp.knn2_l2(dq, dt, mode=E.L2_SIMT)
    torch.cuda.synchronize()
    same_i = bool(torch.equal(ia, ib))
    same_d = bool(torch.equal(da, db))
    nbad = int((ia != ib).any(1).sum().item())
    print("%-34s nq %6d nt %7d  idx %s dist %s  mismatching rows %d  fallback rows %d" % (name, q.shape[0], t.shape[0], same_i, same_d, nbad, fb),
          flush=True)
    return same_i and same_d


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    rng = np.random.default_rng(5)
    ok = True
    for nq, nt in [(300, 900), (1, 5), (129, 3), (257, 64), (256, 65), (4000, 70001), (20000, 1085), (700, 20000)]:
        q = rng.standard_normal((nq, 128)).astype(np.float32)
        t = rng.standard_normal((nt, 128)).astype(np.float32)
        k = min(nq, nt) // 3
        if k:
            t[:k] = q[:k] + 0.05 * rng.standard_normal((k, 128)).astype(np.float32)
        ok &= check("gaussian", q, t)
    q, t = hp.synth.sift_like(3000, 20000, seed=5, planted=0.3)
    ok &= check("sift-like x 1.0137", (q * np.float32(1.0137)).astype(np.float32), (t * np.float32(1.0137)).astype(np.float32))
    q, t = hp.synth.sift_like(5000, 200000, seed=6, planted=0.3)
    ok &= check("sift-like x 1.0137 (200k)", (q * np.float32(1.0137)).astype(np.float32), (t * np.float32(1.0137)).astype(np.float32))
    for scale in (1e-12, 1e12):
        q = (rng.standard_normal((500, 128)) * scale).astype(np.float32)
        t = (rng.standard_normal((3000, 128)) * scale).astype(np.float32)
        ok &= check("gaussian x %g" % scale, q, t)
    q = rng.standard_normal((300, 128)).astype(np.float32)
    t = rng.standard_normal((5000, 128)).astype(np.float32)
    t[:, 0] += 1000.0   # one huge component: coarse scale for everything else
    ok &= check("one dominant component", q, t)
    ok &= check("all zero", np.zeros((130, 128), np.float32), np.zeros((700, 128), np.float32))
    ok &= check("constant rows (flat data)", np.full((130, 128), 3.25, np.float32), np.full((700, 128), 3.25, np.float32) + rng.integers(0, 2, (700, 128)).astype(np.float32))
    q = rng.standard_normal((257, 128)).astype(np.float32)
    t = rng.standard_normal((6000, 128)).astype(np.float32)
    for r in range(40):
        t[100 * r:100 * r + 20] = q[r] + 1e-4 * rng.standard_normal((20, 128)).astype(np.float32)
    ok &= check("near ties", q, t)
    print("PARITY", "OK" if ok else "FAILED", flush=True)

    sizes = [(65536, 262144)]
    if "--big" in sys.argv:
        sizes.append((1 << 20, 1 << 20))
    for nq, nt in sizes:
        q, t = hp.synth.sift_like(nq, nt, seed=1, planted=0.1)
        dq = d(q) * 1.0137
        dt = d(t) * 1.0137
        del q, t
        for name, mode in [("q16", E.L2_TC_Q16), ("split", E.L2_TC_SPLIT)]:
            ws = E.knn2_l2_workspace(nq, nt, 128, mode, dev)
            idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
            dst = torch.empty((nq, 2), dtype=torch.float32, device=dev)
            ms = timed(lambda: E.knn2_l2(dq, dt, mode=mode, out=(idx, dst), workspace=ws), reps=2 if nq > 100000 else 3)
            fb = E.knn2_l2_last_fallback_rows()
            print("%-6s %8d x %8d  %.2f ms  %.3e pairs/s  fallback rows %d" % (name, nq, nt, ms, nq * nt / ms * 1e3, fb), flush=True)
            if name == "q16":
                ref = (idx.clone(), dst.clone())
            else:
                print("   q16 == split: idx %s dist %s" % (bool(torch.equal(ref[0], idx)), bool(torch.equal(ref[1], dst))), flush=True)
            del ws


if __name__ == "__main__":
    main()
