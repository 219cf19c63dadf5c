"""Diagnostic: which database columns does the Q16 engine lose?"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
E = hp.engine
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
for nq, nt in [(256, 64), (256, 128), (256, 256), (512, 1024), (300, 900)]:
    q = rng.standard_normal((nq, 128)).astype(np.float32)
    t = rng.standard_normal((nt, 128)).astype(np.float32)
    dq, dt = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
    ia, da = hp.knn2_l2(dq, dt, mode=E.L2_TC_Q16)
    ib, db = hp.knn2_l2(dq, dt, mode=E.L2_SIMT)
    ia, ib = ia.cpu().numpy(), ib.cpu().numpy()
    bad = np.nonzero((ia != ib).any(1))[0]
    missed = []
    for r in bad:
        for c in ib[r]:
            if c not in ia[r]:
                missed.append(int(c))
    print("nq %d nt %d: %d bad rows; rows: %s" % (nq, nt, len(bad), bad[:40].tolist()))
    print("   missed true-neighbour columns (sorted, unique): %s" % sorted(set(missed)))
    print("   bad rows mod 128 histogram by quarter:", np.bincount((bad % 128) // 32, minlength=4).tolist(), " by m:", np.bincount((bad % 256) // 128, minlength=2).tolist())
