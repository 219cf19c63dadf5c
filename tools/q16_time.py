"""Kernel time of the Q16 engine under its experiment switches (HPUSM_Q16_DBG) at a given size.
    python tools/q16_time.py [--engine i8] [n] [dbg values...]
Needs a development build of the library: HPUSM_EXTRA_FLAGS=-DHPUSM_DEV_SWITCHES bash humanpose-and-urbanscene-matching_b200/csrc/build.sh
(the product build ignores HPUSM_Q16_DBG / HPUSM_I8_DBG: every variant but 0 gives WRONG answers).
"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
E = hp.engine
dev = torch.device("cuda:0")
engine = "q16"
if "--engine" in sys.argv:
    k = sys.argv.index("--engine")
    engine = sys.argv[k + 1]
    del sys.argv[k:k + 2]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
variants = sys.argv[2:] or ["0", "1", "2"]
MODE = E.L2_TC_Q16 if engine == "q16" else E.L2_TC_I8
ENV = "HPUSM_Q16_DBG" if engine == "q16" else "HPUSM_I8_DBG"
SCALE = 1.0137 if engine == "q16" else 1.0
TN, MMAS = (64, 13) if engine == "q16" else (128, 5)
q, t = hp.synth.sift_like(n, n, seed=1, planted=0.1)
dq = torch.from_numpy(q).to(dev) * SCALE
dt = torch.from_numpy(t).to(dev) * SCALE
ws = E.knn2_l2_workspace(n, n, 128, MODE, dev)
idx = torch.empty((n, 2), dtype=torch.int32, device=dev)
dst = torch.empty((n, 2), dtype=torch.float32, device=dev)
E.profile_enable(True)
for v in variants:
    os.environ[ENV] = v
    for _ in range(2):
        E.knn2_l2(dq, dt, mode=MODE, out=(idx, dst), workspace=ws)
    torch.cuda.synchronize()
    E.profile_collect()
    reps = 5
    for _ in range(reps):
        E.knn2_l2(dq, dt, mode=MODE, out=(idx, dst), workspace=ws)
    torch.cuda.synchronize()
    ms, k = E.profile_collect()
    ms /= max(k, 1)
    ns_per_piece = ms * 1e6 / ((n / 128) * (n / TN) / 148)
    print("%s dbg %s: kernel %.3f ms  %.3e pairs/s  %.0f ns per 128 x %d piece (%d MMAs)  fallback %d" %
          (engine, v, ms, n * n / ms * 1e3, ns_per_piece, TN, MMAS, E.knn2_l2_last_fallback_rows()), flush=True)
