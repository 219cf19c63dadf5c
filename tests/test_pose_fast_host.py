"""The FP32 pose fast path (csrc/pose_fast.cuh, __host__ __device__) run on the CPU against a double-precision evaluation
of the same formulas and against the oracle (oracle/pose_oracle.py = the reference's match_single restated).
What must hold: a pose the fast path CERTIFIES has the double path's decision -- never a mismatch -- and only a small
fraction of the poses is left to the exact kernel.  CPU only."""
import ctypes
import importlib
import os
import subprocess

import numpy as np
import pytest

from oracle import pose_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THR = np.array([0.236, 19.8, 0.082, 24.0, 0.125], np.float64)  # thresholds.py:7-13


@pytest.fixture(scope="module")
def fast(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pf") / "libposefast.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "pose_fast_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    P = ctypes.c_void_p
    lib.pose_fast_host.argtypes = [P, P, ctypes.c_int64, ctypes.c_int, P, ctypes.c_int, P, P, P]

    def run(q, db, qim, precise):
        q = np.ascontiguousarray(q, np.float32)
        db = np.ascontiguousarray(db, np.float32)
        n = len(db)
        m, s, c = np.zeros(n, np.uint8), np.zeros(n, np.float64), np.zeros(n, np.uint8)
        usable = lib.pose_fast_host(q.ctypes.data, db.ctypes.data, n, qim, THR.ctypes.data, precise, m.ctypes.data, s.ctypes.data,
                                    c.ctypes.data)
        return m.astype(bool), s, c.astype(bool), usable
    return run


@pytest.mark.parametrize("qim", [1, 0])
def test_certified_decisions_are_the_double_path_decisions(fast, qim):
    synth = importlib.import_module("humanpose-and-urbanscene-matching_b200.synth")
    q = synth.CANONICAL_POSE.astype(np.float32)
    db = synth.pose_database(synth.CANONICAL_POSE, 400000, seed=11 + qim)
    mf, sf, cf, usable = fast(q, db, qim, 0)
    md, sd, cd, _ = fast(q, db, qim, 1)
    assert usable == 1
    both = cf & cd
    assert not (both & (mf != md)).any()                       # never a wrong certified decision
    assert np.abs(sf[both] - sd[both]).max() < 2e-6            # certified scores agree with the double evaluation
    assert cf.mean() > 0.97 and not (cf & ~cd).any()           # ~1 % of the poses go to the exact kernel
    # the double evaluation of these formulas IS the reference's answer (oracle = match_single restated in NumPy)
    k = 1500
    om, osc = pose_oracle.match_single_batch(q.astype(np.float64), db[:k].astype(np.float64), bool(qim))
    sel = cd[:k]
    np.testing.assert_array_equal(md[:k][sel], om[sel])
    np.testing.assert_allclose(sd[:k][sel], osc[sel], rtol=0, atol=1e-12)


def test_poses_near_the_thresholds_are_deferred_not_misjudged(fast):
    """A database built to sit on the decision boundaries: scaled copies of the query with the one joint that sets the
    torso distance moved across the threshold in steps of 1e-7 of the normalised frame."""
    synth = importlib.import_module("humanpose-and-urbanscene-matching_b200.synth")
    q = synth.CANONICAL_POSE.astype(np.float32)
    base = synth.CANONICAL_POSE.copy()
    rng = np.random.default_rng(3)
    db = np.repeat(base[None], 20000, 0)
    db[:, 4, 0] += np.linspace(0, 120, 20000) + rng.normal(0, 1e-3, 20000)   # right wrist sweeps outwards
    mf, sf, cf, _ = fast(q, db, 1, 0)
    md, sd, cd, _ = fast(q, db, 1, 1)
    both = cf & cd
    assert not (both & (mf != md)).any()
    assert md.any() and (~md).any()                            # the sweep crosses the boundary
    assert (~cf).sum() >= 1                                    # and the poses on it were handed to the exact kernel


def test_fast_path_declines_what_it_cannot_do(fast):
    synth = importlib.import_module("humanpose-and-urbanscene-matching_b200.synth")
    q = synth.CANONICAL_POSE.astype(np.float32)
    db = synth.pose_database(synth.CANONICAL_POSE, 64, seed=5)
    db[0] = 0                      # no joint detected
    db[1, :, 0] = 100.0            # zero extent in x
    db[2, 3] = [-5.0, 40.0]        # negative coordinate: the reference's clamp path
    db[3, 2:] = 0                  # only two joints left
    mf, sf, cf, _ = fast(q, db, 1, 0)
    assert not cf[:4].any()
    qn = q.copy()
    qn[5, 0] = -1.0                # a query with a negative coordinate: the whole launch is the exact kernel's
    assert fast(qn, db, 1, 0)[3] == 0
