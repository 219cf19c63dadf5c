"""Kernel time of the Q16 engine under its experiment switches (HPUSM_Q16_DBG) at a given size.
    python tools/q16_time.py [n] [dbg values...]
"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
E = hp.engine
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
variants = sys.argv[2:] or ["0", "1", "2"]
q, t = hp.synth.sift_like(n, n, seed=1, planted=0.1)
dq = torch.from_numpy(q).to(dev) * 1.0137
dt = torch.from_numpy(t).to(dev) * 1.0137
ws = E.knn2_l2_workspace(n, n, 128, E.L2_TC_Q16, dev)
idx = torch.empty((n, 2), dtype=torch.int32, device=dev)
dst = torch.empty((n, 2), dtype=torch.float32, device=dev)
E.profile_enable(True)
for v in variants:
    os.environ["HPUSM_Q16_DBG"] = v
    for _ in range(2):
        E.knn2_l2(dq, dt, mode=E.L2_TC_Q16, out=(idx, dst), workspace=ws)
    torch.cuda.synchronize()
    E.profile_collect()
    reps = 5
    for _ in range(reps):
        E.knn2_l2(dq, dt, mode=E.L2_TC_Q16, out=(idx, dst), workspace=ws)
    torch.cuda.synchronize()
    ms, k = E.profile_collect()
    ms /= max(k, 1)
    cyc_per_piece = ms * 1e-3 * 1.85e9 / ((n / 128) * (n / 64) / 148)
    print("dbg %s: kernel %.3f ms  %.3e pairs/s  ~%.0f cycles per piece at 1.85 GHz (13 MMAs = 416)  fallback %d" %
          (v, ms, n * n / ms * 1e3, cyc_per_piece, E.knn2_l2_last_fallback_rows()), flush=True)
