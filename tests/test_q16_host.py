"""CPU checks of the arithmetic behind the fixed-point integer engine for general float descriptors (l2_tc_q16.cu),
through its NumPy restatement oracle/q16_oracle.py: the error bound E(q) the proof of completeness rests on really bounds
|K - 256 delta^2 key| against float64 truth on adversarial data, and rows the proof accepts carry the exact answer.
(The CUDA kernels are held against the exact engine on the GPU in tests/test_gpu_parity.py::test_l2_q16_*.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import q16_oracle  # noqa: E402


def _sift_like(rng, n):
    x = rng.gamma(0.6, 1.0, (n, 128))
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    return (np.minimum(np.rint(512.0 * x), 255.0) * 1.0137).astype(np.float32)


def _datasets():
    rng = np.random.default_rng(7)
    nq, nt = 120, 1500
    out = {}
    out["gaussian"] = (rng.standard_normal((nq, 128)).astype(np.float32), rng.standard_normal((nt, 128)).astype(np.float32))
    out["sift-like x 1.0137"] = (_sift_like(rng, nq), _sift_like(rng, nt))
    out["gaussian x 1e-9"] = tuple((a * 1e-9).astype(np.float32) for a in out["gaussian"])
    out["gaussian x 1e9"] = tuple((a * 1e9).astype(np.float32) for a in out["gaussian"])
    t = rng.standard_normal((nt, 128)).astype(np.float32)
    t[:, 0] += 1000.0
    out["one dominant component"] = (rng.standard_normal((nq, 128)).astype(np.float32), t)
    flat = np.full((nt, 128), 3.25, np.float32) + rng.integers(0, 2, (nt, 128)).astype(np.float32)
    out["flat rows (norm constraint sets the scale)"] = (np.full((nq, 128), 3.25, np.float32) + rng.random((nq, 128)).astype(np.float32), flat)
    heavy = np.clip(rng.standard_cauchy((nt, 128)), -50, 50).astype(np.float32)
    out["heavy tails"] = (np.clip(rng.standard_cauchy((nq, 128)), -50, 50).astype(np.float32), heavy)
    # values sitting just below quantisation half-steps: every component rounds with the largest possible residue
    q, t = out["gaussian"]
    delta, _ = q16_oracle.scale(q, t)
    half = lambda a: ((np.floor(a / delta) + 0.4999) * delta).astype(np.float32)  # noqa: E731
    out["half-step residues"] = (half(q), half(t))
    out["mixed signs, sparse"] = ((q * (rng.random(q.shape) < 0.2)).astype(np.float32), (t * (rng.random(t.shape) < 0.2)).astype(np.float32))
    return out


@pytest.mark.parametrize("name", list(_datasets().keys()))
def test_q16_key_error_bound_holds(name):
    q, t = _datasets()[name]
    key, E, delta = q16_oracle.keys_and_bound(q, t)
    q64, t64 = q.astype(np.float64), t.astype(np.float64)
    K = 2.0 * q64 @ t64.T - (t64 ** 2).sum(axis=1)[None, :]
    err = np.abs(K - 256.0 * np.float64(delta) ** 2 * key.astype(np.float64))
    worst = (err / E[:, None]).max()
    assert worst <= 1.0, (name, worst)
    assert worst > 0.05, (name, worst)   # ... and is not vacuous: on these data the error reaches 15-31 % of it


def test_q16_proven_rows_carry_the_exact_answer():
    rng = np.random.default_rng(8)
    for name, (q, t) in _datasets().items():
        if "flat" in name or "half-step" in name:
            continue
        q, t = q[:60], t[:900].copy()
        t[:20] = q[:20] + (0.05 * np.abs(t).max() / 5.0) * rng.standard_normal((20, 128)).astype(np.float32)  # planted neighbours
        idx, proven = q16_oracle.knn2_with_proof(q, t)
        d2 = ((q[:, None, :].astype(np.float64) - t[None, :, :].astype(np.float64)) ** 2).sum(-1)
        truth = np.argsort(d2, axis=1, kind="stable")[:, :2]
        gap_ok = np.sort(d2, axis=1)[:, 2] - np.sort(d2, axis=1)[:, 1] > 1e-5 * np.sort(d2, axis=1)[:, 1]  # no near-tie at rank 2/3
        sel = proven & gap_ok
        np.testing.assert_array_equal(idx[sel], truth[sel], err_msg=name)
        # a component 200 x the others sets a coarse scale for all of them: the proof covers fewer rows there (the rest take
        # the exact engine), everywhere else it covers (nearly) every row
        assert proven.mean() > (0.5 if "dominant" in name else 0.9), (name, proven.mean())


def test_q16_near_ties_are_not_proven():
    """Seven database rows within the bound of each other around a query: a 6-deep list cannot be proven complete."""
    rng = np.random.default_rng(9)
    q = rng.standard_normal((8, 128)).astype(np.float32)
    t = rng.standard_normal((640, 128)).astype(np.float32)
    for r in range(8):
        t[64 * r:64 * r + 7] = q[r] + 1e-4 * rng.standard_normal((7, 128)).astype(np.float32)   # all in one 32-column list
    _, proven = q16_oracle.knn2_with_proof(q, t)
    assert not proven.any()
