"""ctypes wrapper of oracle/ransac_oracle.c -- test infrastructure only (see the header of the C file)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libransac_oracle.so")

SAMPLER_COUNTER = 0  # counter-based table keyed on (seed, pair, hypothesis, attempt)
SAMPLER_CV2 = 1      # OpenCV's own fixed-seed sequence: reproduces cv2.findHomography


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
        p, i64, u64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
        _lib.ransac_oracle_score.argtypes = [p, p, p, i64, ctypes.c_float, i32, i32, u64, i64, p]
        _lib.ransac_oracle_homography.argtypes = [p, p, p, i64, ctypes.c_float, i32, ctypes.c_double, i32, u64, i64, i32,
                                                  p, p, p, p, p]
        _lib.ransac_oracle_refine.argtypes = [p, p, p, i64, p]
        _lib.ransac_oracle_sample4.argtypes = [u64, u64, u64, u32, u32, p]
        _lib.ransac_oracle_samples.argtypes = [p, p, u32, i32, u64, u64, i32, p]
        _lib.ransac_oracle_hypothesis.argtypes = [p, p, u32, i32, u64, u64, i32, p]
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def sample4(seed, pair, hyp, attempt, K):
    out = np.zeros(4, np.uint32)
    _load().ransac_oracle_sample4(seed, pair, hyp, attempt, K, _p(out))
    return out


def samples(src, dst, nhyp, sampler=SAMPLER_COUNTER, seed=0, pair=0):
    """The accepted 4-point sample of every iteration of one pair: int32 [nhyp, 4], -1 where there is none."""
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    out = np.empty((nhyp, 4), np.int32)
    _load().ransac_oracle_samples(_p(src), _p(dst), src.shape[0], sampler, seed, pair, nhyp, _p(out))
    return out


def hypothesis(src, dst, seed, pair, hyp, sampler=SAMPLER_COUNTER):
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    H = np.zeros(9, np.float64)
    ok = _load().ransac_oracle_hypothesis(_p(src), _p(dst), src.shape[0], sampler, seed, pair, hyp, _p(H))
    return (H.reshape(3, 3) if ok else None)


def score_batch(src, dst, pair_offsets, thresh=5.0, nhyp=1024, seed=0, sampler=SAMPLER_COUNTER, pair_index0=0):
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    offs = _c(pair_offsets, np.int64)
    npairs = offs.size - 1
    counts = np.empty((npairs, nhyp), np.int32)
    _load().ransac_oracle_score(_p(src), _p(dst), _p(offs), npairs, thresh, nhyp, sampler, seed, pair_index0, _p(counts))
    return counts


def homography_batch(src, dst, pair_offsets, thresh=5.0, nhyp=2000, confidence=0.995, seed=0, refine=True,
                     sampler=SAMPLER_COUNTER, pair_index0=0, return_min_mask=False):
    """Same contract as hpusm_ransac_homography_batch: (H [npairs,3,3], mask [total], ninl, best_hyp)."""
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    offs = _c(pair_offsets, np.int64)
    npairs = offs.size - 1
    H = np.empty((npairs, 3, 3), np.float64)
    mask = np.zeros(src.shape[0], np.uint8)
    min_mask = np.zeros(src.shape[0], np.uint8)
    ninl = np.empty(npairs, np.int32)
    best = np.empty(npairs, np.int32)
    _load().ransac_oracle_homography(_p(src), _p(dst), _p(offs), npairs, thresh, nhyp, confidence, sampler, seed,
                                     pair_index0, 1 if refine else 0, _p(H), _p(mask), _p(ninl), _p(best), _p(min_mask))
    if return_min_mask:
        return H, mask, ninl, best, min_mask
    return H, mask, ninl, best


def find_homography_cv2(src, dst, thresh=5.0, max_iters=2000, confidence=0.995):
    """cv2.findHomography(src, dst, cv2.RANSAC, thresh) restated: (H or None, mask uint8 (K,1))."""
    src = _c(src, np.float32).reshape(-1, 2)
    offs = np.array([0, src.shape[0]], np.int64)
    H, mask, ninl, best = homography_batch(src, dst, offs, thresh, max_iters, confidence, 0, True, SAMPLER_CV2)
    if best[0] < 0:
        return None, mask.reshape(-1, 1)
    return H[0], mask.reshape(-1, 1)


def refine(src, dst, mask):
    """runKernel (normalised DLT) + LM(10) homography on the rows flagged in mask (cv2's post-RANSAC step)."""
    src, dst = _c(src, np.float32).reshape(-1, 2), _c(dst, np.float32).reshape(-1, 2)
    mask = _c(mask, np.uint8).ravel()
    H = np.empty((3, 3), np.float64)
    rc = _load().ransac_oracle_refine(_p(src), _p(dst), _p(mask), src.shape[0], _p(H))
    if rc != 0:
        raise ValueError("fewer than 4 inliers or a degenerate point set")
    return H
