import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
E = hp.engine
dev = torch.device("cuda:0")
from test_gpu_fullsize import _sift_like
n = 1 << 20
t = _sift_like(n, 1, dev); q = _sift_like(n, 2, dev)
g = torch.Generator(device=dev).manual_seed(3)
planted_q = torch.randperm(n, generator=g, device=dev)[:100000]
planted_t = torch.randint(0, n, (100000,), generator=g, device=dev)
q[planted_q] = torch.clamp(t[planted_t] + torch.randn((100000, 128), generator=g, device=dev) * 2.0, 0, 255)
q *= 1.0137; t *= 1.0137
idx, dist = hp.knn2_l2(q, t); print("whole: mode", E.knn2_l2_last_mode(), "fallback", E.knn2_l2_last_fallback_rows())
h = n // 2
ia, da = hp.knn2_l2(q, t[:h].contiguous()); print("half a fallback", E.knn2_l2_last_fallback_rows())
ib, db = hp.knn2_l2(q, t[h:].contiguous()); print("half b fallback", E.knn2_l2_last_fallback_rows())
mi, md = E.knn2_merge(torch.stack([ia, torch.where(ib >= 0, ib + h, ib)]), torch.stack([da, db]))
bad = torch.nonzero((mi != idx).any(1) | (md != dist).any(1)).flatten()
print("mismatching rows:", bad.numel(), bad[:20].tolist())
rows = bad[:2048]
if rows.numel():
    qs = q[rows].contiguous()
    xi, xd = hp.knn2_l2(qs, t, mode=E.L2_SIMT)
    print("whole == exact on those rows:", int(((xi == idx[rows]).all(1) & (xd == dist[rows]).all(1)).sum()), "of", rows.numel())
    print("merged == exact on those rows:", int(((xi == mi[rows]).all(1) & (xd == md[rows]).all(1)).sum()), "of", rows.numel())
    for r in range(min(5, rows.numel())):
        print(" row", int(rows[r]), "whole", idx[rows[r]].tolist(), dist[rows[r]].tolist(), "merged", mi[rows[r]].tolist(), md[rows[r]].tolist(), "exact", xi[r].tolist(), xd[r].tolist())
