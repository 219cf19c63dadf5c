"""TEST INFRASTRUCTURE ONLY (never imported by the product): NumPy restatement of the arithmetic of the fixed-point
integer engine for general float descriptors (humanpose-and-urbanscene-matching_b200/csrc/l2_tc_q16.cu), which replaces
cv2.BFMatcher(cv2.NORM_L2).knnMatch(..., k=2) at reference urbanscene/features.py:59 for non-integer descriptors.

The engine orders database rows by an INTEGER key built from 15-bit fixed-point operands and proves, per query, that
its candidate lists contain the true two nearest rows.  The proof rests on a bound E(q) of |K - 256 delta^2 key| with
K = 2 q.t - ||t||^2.  This module recomputes keys and bound the way the kernels do, so that the CPU suite can check the
bound against float64 truth on adversarial data -- independently of any GPU.
"""
import numpy as np

XMAX = 16256          # 127 * 128
TSCALE = np.float32(1.55e10)
N1_MAX = 480000
M = 6                 # candidates per list (Q_M)
COLS_PER_LIST = 32    # a list = one 32-column group of every 64-row tile


def scale(q, t):
    """delta as l2_q16_params_kernel computes it (float32)."""
    amax = np.float32(max(np.abs(q).max(initial=0.0), np.abs(t).max(initial=0.0)))
    tmax = np.float32((t.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32).max(initial=0.0))
    d = np.float32(max(amax / np.float32(XMAX), np.sqrt(tmax / TSCALE))) * np.float32(1.00001)
    if not (d > 0) or not np.isfinite(d):
        d = np.float32(1.0)
    return np.float32(max(d, np.float32(1e-30))), tmax


def quantise(x, delta):
    """X = 128 H + L and the residue e = x - delta X (q16_quantise)."""
    inv = np.float32(1.0) / delta
    xf = np.rint((x.astype(np.float32) * inv).astype(np.float32))
    xf = np.clip(xf, -XMAX, XMAX)
    e = x.astype(np.float64) - xf.astype(np.float64) * np.float64(delta)
    X = xf.astype(np.int64)
    H = (X + 64) >> 7
    L = X - 128 * H
    assert np.abs(H).max(initial=0) <= 127 and L.min(initial=0) >= -64 and L.max(initial=0) <= 63
    return H, L, e


def keys_and_bound(q, t):
    """(key [nq, nt] int64, E [nq] float64, delta): the kernel's integer key of every pair and the finalize kernel's bound."""
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    delta, tmax = scale(q, t)
    inv = np.float32(1.0) / delta
    Hq, Lq, eq = quantise(q, delta)
    Ht, Lt, et = quantise(t, delta)
    s = (t ** 2).sum(axis=1, dtype=np.float32)
    n1 = np.rint((s * inv * inv * np.float32(1.0 / 32768.0)).astype(np.float32)).astype(np.int64)
    assert n1.max(initial=0) <= N1_MAX
    key = 128 * (Hq @ Ht.T - n1[None, :]) + Hq @ Lt.T + Lq @ Ht.T
    assert np.abs(key).max(initial=0) < 2 ** 30
    dl = np.float64(delta)
    qn = (q.astype(np.float64) ** 2).sum(axis=1)
    rq = np.sqrt((eq ** 2).sum(axis=1))
    lq = dl * np.sqrt((Lq ** 2).sum(axis=1).astype(np.float64))
    rt = np.sqrt((et ** 2).sum(axis=1).max(initial=0.0))
    lt = dl * np.sqrt(np.float64((Lt ** 2).sum(axis=1).max(initial=0)))
    tn = np.sqrt(np.float64(tmax))
    qnorm = np.sqrt(qn)
    E = (1.01 * (2.0 * (qnorm + rq) * rt + 2.0 * (tn + rt) * rq + 2.0 * rq * rt + 2.0 * lq * lt) + 20000.0 * dl * dl +
         4e-5 * (qn + np.float64(tmax)) + 4e-5 * qnorm * tn)
    return key, E, delta


def knn2_with_proof(q, t):
    """The engine's answer for one database split: per query the top-M keys of every 32-column list -> exact float32
    distances of the listed rows -> lexicographic top-2 -> proof.  Returns (idx [nq,2], proven [nq] bool)."""
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    key, E, delta = keys_and_bound(q, t)
    nq, nt = key.shape
    dl = np.float64(delta)
    list_of = (np.arange(nt) // COLS_PER_LIST) % 2          # column group 0 / 1 of the 64-row tile
    idx = np.full((nq, 2), -1, np.int64)
    proven = np.zeros(nq, bool)
    for r in range(nq):
        cands, kmax, complete = [], -np.inf, True
        for g in (0, 1):
            cols = np.nonzero(list_of == g)[0]
            if cols.size == 0:
                continue
            order = cols[np.lexsort((cols, -key[r, cols]))]    # key descending, lower column first among equals
            cands.extend(order[:M].tolist())
            if cols.size > M:
                kmax = max(kmax, float(key[r, order[M - 1]]))
                complete = False
        c = np.array(sorted(cands))
        d2 = ((q[r][None, :] - t[c]) ** 2).sum(axis=1, dtype=np.float32)
        o = np.lexsort((c, d2))
        idx[r, :min(2, c.size)] = c[o[:2]]
        qn = np.float64((q[r].astype(np.float32) ** 2).sum(dtype=np.float32))
        if complete:
            proven[r] = True
        elif c.size >= 2:
            proven[r] = float(d2[o[1]]) + E[r] < qn - 256.0 * dl * dl * kmax
    return idx, proven
