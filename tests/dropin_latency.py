#!/usr/bin/env python
"""(run as a script: python tests/dropin_latency.py; lives under tests/ because it times the CPU oracle beside the product)
Per-call latency of the scalar drop-ins (one 18-joint problem per call: H2D + launch + D2H) next to the same call on the
CPU: the reference's own functions when /root/reference is present (build container), else the NumPy restatement in
oracle/pose_oracle.py (GPU box).  SURVEY section 6 measured the reference at ~400 us (single_person.match_single),
~71 us (common.find_transformation) and ~41 us (procrustes.procrustes) per call.  There is no CPU fallback in the
drop-ins: where a single call loses to the CPU, callers belong on the batched entry points (engine.pose_match_batch,
engine.affine_lsq_batch, engine.procrustes_batch), whose per-problem cost is printed beside it."""
import importlib
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
DROPIN = "humanpose-and-urbanscene-matching_b200.dropin"
sp = importlib.import_module(DROPIN + ".posematching.single_person")
pr = importlib.import_module(DROPIN + ".posematching.procrustes")
cm = importlib.import_module(DROPIN + ".common")
from oracle import pose_oracle  # noqa: E402


def per_call_us(fn, n=200):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


ref = None
if os.path.isdir("/root/reference"):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    sys.path.insert(0, "/root/reference")
    np.math = math
    import common as ref_common
    import posematching.procrustes as ref_procrustes
    import posematching.single_person as ref_sp
    ref = {"match_single": ref_sp.match_single, "find_transformation": ref_common.find_transformation,
           "procrustes": ref_procrustes.procrustes}

db = hp.synth.pose_database(hp.synth.CANONICAL_POSE, 4096, seed=1).astype(np.float64)
q = hp.synth.CANONICAL_POSE.astype(np.float64)
a, b = db[5], db[9]
dev = torch.device("cuda:0")
out = {"cpu_kind": "reference" if ref else "port (oracle/pose_oracle.py)"}
cpu = ref or {"match_single": lambda m, i: pose_oracle.match_single(m, i),
              "find_transformation": pose_oracle.find_transformation, "procrustes": pose_oracle.procrustes}
out["match_single"] = {"dropin_us": per_call_us(lambda: sp.match_single(q, a)), "cpu_us": per_call_us(lambda: cpu["match_single"](q, a))}
out["find_transformation"] = {"dropin_us": per_call_us(lambda: cm.find_transformation(a[:10], b[:10])),
                              "cpu_us": per_call_us(lambda: cpu["find_transformation"](a[:10], b[:10]))}
out["procrustes"] = {"dropin_us": per_call_us(lambda: pr.procrustes(a, b)), "cpu_us": per_call_us(lambda: cpu["procrustes"](a, b))}
# the batched forms the callers are pointed at: per-problem cost over 4096 problems, inputs resident
dq, ddb = torch.from_numpy(q.astype(np.float32)).to(dev), torch.from_numpy(db.astype(np.float32)).to(dev)
X, Y = torch.from_numpy(db).to(dev), torch.from_numpy(np.roll(db, 1, 0)).to(dev)
out["batched_per_problem_us"] = {
    "pose_match_batch": per_call_us(lambda: hp.engine.pose_match_batch(dq, ddb), 50) / len(db),
    "affine_lsq_batch": per_call_us(lambda: hp.engine.affine_lsq_batch(X, Y), 50) / len(db),
    "procrustes_batch": per_call_us(lambda: hp.engine.procrustes_batch(X, Y), 50) / len(db)}
print(json.dumps(out))
