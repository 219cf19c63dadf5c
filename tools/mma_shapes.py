"""Issue rate of back-to-back tcgen05.mma kind::i8 at the kNN engines' shapes."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
dev = torch.device("cuda:0")
for kind in ("i8", "i8n128", "i8n128t", "i8n64", "i8n64t"):
    n = {"i8": 256, "i8n128": 128, "i8n128t": 128, "i8n64": 64, "i8n64t": 64}[kind]
    hp.engine.microbench_mma(kind, 200, dev)
    torch.cuda.synchronize()
    rounds = 4000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops = hp.engine.microbench_mma(kind, rounds, dev); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("mma %-8s: %.2f ms -> %.0f T op/s; %.1f ns per MMA = %.1f cycles at 1.9 GHz (N/2 = %d)" %
          (kind, ms, ops / ms / 1e9, ms * 1e6 / (rounds * 64), ms * 1e6 / (rounds * 64) * 1.9, n // 2))
