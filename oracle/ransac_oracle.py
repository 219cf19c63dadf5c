"""ctypes wrapper of oracle/ransac_oracle.c -- test infrastructure only (see the header of the C file)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libransac_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ransac_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        p, i64, u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64
        _lib.ransac_oracle_score.argtypes = [p, p, p, i64, ctypes.c_float, ctypes.c_int, u64, p]
        _lib.ransac_oracle_homography.argtypes = [p, p, p, i64, ctypes.c_float, ctypes.c_int, ctypes.c_double, u64,
                                                  ctypes.c_int, p, p, p, p]
        _lib.ransac_oracle_refine.argtypes = [p, p, p, i64, p, p]
        _lib.ransac_oracle_sample4.argtypes = [u64, u64, u64, ctypes.c_uint32, p]
        _lib.ransac_oracle_hypothesis.argtypes = [p, p, ctypes.c_uint32, u64, u64, u64, p]
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def sample4(seed, pair, hyp, K):
    out = np.zeros(4, np.uint32)
    _load().ransac_oracle_sample4(seed, pair, hyp, K, _p(out))
    return out


def hypothesis(src, dst, seed, pair, hyp):
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    H = np.zeros(9, np.float64)
    ok = _load().ransac_oracle_hypothesis(_p(src), _p(dst), src.shape[0], seed, pair, hyp, _p(H))
    return (H.reshape(3, 3) if ok else None)


def score_batch(src, dst, pair_offsets, thresh=5.0, nhyp=1024, seed=0):
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    offs = _c(pair_offsets, np.int64)
    npairs = offs.size - 1
    counts = np.empty((npairs, nhyp), np.int32)
    _load().ransac_oracle_score(_p(src), _p(dst), _p(offs), npairs, thresh, nhyp, seed, _p(counts))
    return counts


def homography_batch(src, dst, pair_offsets, thresh=5.0, nhyp=2000, confidence=0.995, seed=0, refine=True):
    """Same contract as hpusm_ransac_homography_batch: (H [npairs,3,3], mask [total], ninl, best_hyp)."""
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    offs = _c(pair_offsets, np.int64)
    npairs = offs.size - 1
    H = np.empty((npairs, 3, 3), np.float64)
    mask = np.zeros(src.shape[0], np.uint8)
    ninl = np.empty(npairs, np.int32)
    best = np.empty(npairs, np.int32)
    _load().ransac_oracle_homography(_p(src), _p(dst), _p(offs), npairs, thresh, nhyp, confidence, seed,
                                     1 if refine else 0, _p(H), _p(mask), _p(ninl), _p(best))
    return H, mask, ninl, best


def refine(src, dst, mask, H_init=None):
    """Least-squares + LM homography on the rows flagged in mask (cv2's post-RANSAC step)."""
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    mask = _c(mask, np.uint8).ravel()
    Hi = _c(np.eye(3) if H_init is None else H_init, np.float64)
    H = np.empty((3, 3), np.float64)
    rc = _load().ransac_oracle_refine(_p(src), _p(dst), _p(mask), src.shape[0], _p(Hi), _p(H))
    if rc != 0:
        raise ValueError("fewer than 4 inliers")
    return H
