"""Brute-force kNN (k=2) + Lowe ratio oracle in NumPy -- test infrastructure only.

Restates what the reference gets from OpenCV at urbanscene/features.py:59
(``matcher.knnMatch(desc_model, trainDescriptors=desc_input, k=2)`` with the matcher of features.py:42)
and the ratio test of features.py:45-55.  OpenCV itself is not in /root/reference; behaviour pinned by
probe and by tests/golden/knn_*.npz recorded from cv2 4.13.0:
  * distances: Hamming = popcount(a ^ b); L2 = sqrt(sum (a-b)^2) as float32
  * order: ascending distance, ties resolved to the LOWER train index (stable)
  * fewer than 2 train rows -> shorter lists (idx -1 here)
"""
import numpy as np

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint8)


def _top2_from_matrix(d):
    """Lexicographic (distance, index) top-2 of every row of d [nq, nt] (nt >= 1)."""
    nq, nt = d.shape
    i0 = np.argmin(d, axis=1)  # first occurrence = lowest index among ties
    rows = np.arange(nq)
    d0 = d[rows, i0]
    if nt < 2:
        return np.stack([i0, -np.ones_like(i0)], 1), np.stack([d0, np.zeros_like(d0)], 1)
    saved = d0.copy()
    big = np.iinfo(d.dtype).max if np.issubdtype(d.dtype, np.integer) else np.inf
    d[rows, i0] = big
    i1 = np.argmin(d, axis=1)
    d1 = d[rows, i1]
    d[rows, i0] = saved
    return np.stack([i0, i1], 1), np.stack([d0, d1], 1)


def _merge_top2(idx_a, d_a, idx_b, d_b):
    """Merge two top-2 lists per row; list a holds LOWER indices than list b (so a wins ties)."""
    nq = idx_a.shape[0]
    idx = np.concatenate([idx_a, idx_b], 1)
    d = np.concatenate([d_a, d_b], 1).astype(np.float64)
    d[idx < 0] = np.inf
    order = np.argsort(d, axis=1, kind="stable")[:, :2]  # stable: earlier column (lower index) wins ties
    rows = np.arange(nq)[:, None]
    return idx[rows, order], np.concatenate([d_a, d_b], 1)[rows, order]


def knn2_hamming(q, t, chunk=256):
    """idx [nq,2] int32 (-1 = missing), dist [nq,2] int32."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    nq, nt = q.shape[0], t.shape[0]
    idx = -np.ones((nq, 2), np.int32)
    dist = np.zeros((nq, 2), np.int32)
    if nq == 0 or nt == 0:
        return idx, dist
    for a in range(0, nq, chunk):
        qa = q[a:a + chunk]
        d = np.zeros((qa.shape[0], nt), np.int32)
        for b in range(q.shape[1]):
            d += _POP8[qa[:, b][:, None] ^ t[:, b][None, :]]
        i2, d2 = _top2_from_matrix(d)
        idx[a:a + chunk] = i2
        dist[a:a + chunk] = np.where(i2 >= 0, d2, 0)
    return idx, dist


def knn2_l2(q, t, chunk=512, tchunk=65536):
    """idx [nq,2] int32, dist [nq,2] float32 = sqrt of the float64-exact squared distance.

    The squared distances are formed in float64 as |a|^2 + |b|^2 - 2ab only when every value is an
    integer below 2**20 (then all terms are exact); otherwise as sum((a-b)^2) in float64."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    nq, nt = q.shape[0], t.shape[0]
    idx = -np.ones((nq, 2), np.int32)
    dist = np.zeros((nq, 2), np.float32)
    if nq == 0 or nt == 0:
        return idx, dist
    q64, t64 = q.astype(np.float64), t.astype(np.float64)
    exact_int = bool(np.all(q64 == np.rint(q64)) and np.all(t64 == np.rint(t64)) and
                     np.abs(q64).max(initial=0) < 2 ** 20 and np.abs(t64).max(initial=0) < 2 ** 20)
    tn = (t64 * t64).sum(1)
    for a in range(0, nq, chunk):
        qa = q64[a:a + chunk]
        best_i = -np.ones((qa.shape[0], 2), np.int64)
        best_d = np.full((qa.shape[0], 2), np.inf)
        for b in range(0, nt, tchunk):
            tb = t64[b:b + tchunk]
            d = (qa * qa).sum(1)[:, None] + tn[None, b:b + tchunk] - 2.0 * (qa @ tb.T)
            if not exact_int:
                # float64 GEMM form (error ~1e-13 |q||t|) to find the few smallest entries, then the exact
                # sum((a-b)^2) in float64 for those, so near-ties are ordered by the true distances
                m = min(4, d.shape[1])
                cols = np.argpartition(d, m - 1, axis=1)[:, :m]
                rows = np.arange(d.shape[0])[:, None]
                diff = tb[cols] - qa[:, None, :]
                exact = np.einsum("ijk,ijk->ij", diff, diff)
                d = np.full_like(d, np.inf)
                d[rows, cols] = exact
            i2, d2 = _top2_from_matrix(d)
            i2 = np.where(i2 >= 0, i2 + b, -1)
            best_i, best_d = _merge_top2(best_i, best_d, i2, d2)
        idx[a:a + chunk] = best_i
        with np.errstate(invalid="ignore"):
            dist[a:a + chunk] = np.where(best_i >= 0, np.sqrt(best_d.astype(np.float32)), 0).astype(np.float32)
    return idx, dist


def ratio_keep(idx, dist, ratio=0.8):
    """features.py:48: len(m) == 2 and m[0].distance < m[1].distance * ratio, evaluated in Python double."""
    d = np.asarray(dist, dtype=np.float64)
    return (np.asarray(idx)[:, 1] >= 0) & (d[:, 0] < d[:, 1] * float(ratio))


def filter_matches_points(kp_q, kp_t, idx, dist, ratio=0.8):
    """features.filter_matches (features.py:45-55) on coordinate arrays: p1, p2 float32 (K,1,2), selected rows."""
    keep = ratio_keep(idx, dist, ratio)
    sel = np.nonzero(keep)[0]
    p1 = np.float32(np.asarray(kp_q)[sel]).reshape(-1, 1, 2)
    p2 = np.float32(np.asarray(kp_t)[np.asarray(idx)[sel, 0]]).reshape(-1, 1, 2)
    return p1, p2, sel


def segmented_scores(q, t, seg_offsets, ratio=0.8):
    """Per database image: number of ratio-test survivors of kNN-2 restricted to the image."""
    seg_offsets = np.asarray(seg_offsets)
    nseg = len(seg_offsets) - 1
    score = np.zeros(nseg, np.int32)
    best = -np.ones((nseg, q.shape[0]), np.int32)
    for s in range(nseg):
        lo, hi = int(seg_offsets[s]), int(seg_offsets[s + 1])
        i2, d2 = knn2_hamming(q, t[lo:hi])
        keep = ratio_keep(i2, d2, ratio)
        score[s] = int(keep.sum())
        best[s] = np.where(keep, i2[:, 0] + lo, -1)
    return score, best
