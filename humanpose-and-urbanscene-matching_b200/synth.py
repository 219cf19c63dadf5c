"""Seeded synthetic workloads of the shapes BASELINE.json names. NumPy only.

C2  sift_like      128-d SIFT-like descriptors: integer valued 0..255 stored as float32, row norm ~512
C3  orb_like       256-bit binary descriptors, planted near-duplicates
C3  orb_database   an image database: nimg images x per_image descriptors + query image
C4  ransac_pair    matches under a random well-conditioned homography with outliers
C5  pose_database  OpenPose 18-joint poses: similarity-transformed, noised fixture poses, some joints undetected
"""
import numpy as np


def _sift_rows(rng, n, d=128):
    x = rng.gamma(0.6, 1.0, (n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    np.minimum(x, 0.2, out=x)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    return np.minimum(np.rint(512.0 * x), 255.0).astype(np.float32)


def sift_like(nq, nt, seed=0, planted=0.1, d=128):
    """(q [nq,d], t [nt,d]) float32, integer valued; `planted` of the queries are noisy copies of db rows."""
    rng = np.random.default_rng(seed)
    t = _sift_rows(rng, nt, d)
    q = _sift_rows(rng, nq, d)
    k = int(round(planted * nq))
    if k and nt:
        qi = rng.choice(nq, k, replace=False)
        ti = rng.integers(0, nt, k)
        noisy = t[ti] + np.rint(rng.normal(0, 4.0, (k, d))).astype(np.float32)
        q[qi] = np.clip(noisy, 0, 255)
    return q, t


def orb_like(nq, nt, seed=1, planted=0.05, nbytes=32, max_flips=24):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (nt, nbytes), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, nbytes), dtype=np.uint8)
    k = int(round(planted * nq))
    if k and nt:
        qi = rng.choice(nq, k, replace=False)
        ti = rng.integers(0, nt, k)
        bits = np.unpackbits(t[ti], axis=1)
        for r in range(k):
            flips = rng.choice(nbytes * 8, rng.integers(0, max_flips + 1), replace=False)
            bits[r, flips] ^= 1
        q[qi] = np.packbits(bits, axis=1)
    return q, t


def orb_database(nimg, per_image=2048, seed=1, planted_images=0.05, nbytes=32, ragged=False):
    """(query [per_image,nbytes], db [total,nbytes], seg_offsets [nimg+1] int64, planted image ids)."""
    rng = np.random.default_rng(seed)
    sizes = np.full(nimg, per_image, np.int64)
    if ragged:
        sizes = rng.integers(max(1, per_image // 4), per_image + 1, nimg).astype(np.int64)
        sizes[rng.integers(0, nimg, max(1, nimg // 16))] = rng.integers(0, 3)  # nearly empty images
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    db = rng.integers(0, 256, (int(offs[-1]), nbytes), dtype=np.uint8)
    query = rng.integers(0, 256, (per_image, nbytes), dtype=np.uint8)
    nplant = max(1, int(round(planted_images * nimg)))
    planted = np.sort(rng.choice(nimg, nplant, replace=False))
    for im in planted:
        n = int(sizes[im])
        m = min(n, per_image) // 2
        if m == 0:
            continue
        rows = rng.choice(n, m, replace=False)
        src = rng.choice(per_image, m, replace=False)
        bits = np.unpackbits(query[src], axis=1)
        flip = rng.random(bits.shape) < (12.0 / (nbytes * 8))
        db[offs[im] + rows] = np.packbits(bits ^ flip.astype(np.uint8), axis=1)
    return query, db, offs, planted


def random_homography(rng):
    th = rng.uniform(-np.pi / 6, np.pi / 6)
    s = rng.uniform(0.7, 1.4)
    H = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-40, 40)],
                  [s * np.sin(th), s * np.cos(th), rng.uniform(-40, 40)],
                  [rng.uniform(-1e-3, 1e-3) * 0.3, rng.uniform(-1e-3, 1e-3) * 0.3, 1.0]])
    return H


def ransac_pair(n, inlier_ratio=0.3, seed=2, noise=1.0, w=640.0, h=480.0):
    """(src [n,2] f32, dst [n,2] f32, H_true): inliers follow H_true + N(0, noise), outliers are uniform."""
    rng = np.random.default_rng(seed)
    H = random_homography(rng)
    src = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1)
    p = np.c_[src, np.ones(n)] @ H.T
    dst = p[:, :2] / p[:, 2:]
    inl = rng.random(n) < inlier_ratio
    dst[inl] += rng.normal(0, noise, (int(inl.sum()), 2))
    dst[~inl] = np.stack([rng.uniform(0, w, int((~inl).sum())), rng.uniform(0, h, int((~inl).sum()))], 1)
    return src.astype(np.float32), dst.astype(np.float32), H


def ransac_batch(npairs, n, inlier_ratio=0.3, seed=2):
    """(src [npairs*n,2], dst, pair_offsets [npairs+1] int64), pair p seeded with seed+p (vectorised)."""
    rng = np.random.default_rng(seed)
    th = rng.uniform(-np.pi / 6, np.pi / 6, npairs)
    s = rng.uniform(0.7, 1.4, npairs)
    H = np.zeros((npairs, 3, 3))
    H[:, 0, 0] = s * np.cos(th); H[:, 0, 1] = -s * np.sin(th); H[:, 0, 2] = rng.uniform(-40, 40, npairs)
    H[:, 1, 0] = s * np.sin(th); H[:, 1, 1] = s * np.cos(th); H[:, 1, 2] = rng.uniform(-40, 40, npairs)
    H[:, 2, 0] = rng.uniform(-3e-4, 3e-4, npairs); H[:, 2, 1] = rng.uniform(-3e-4, 3e-4, npairs); H[:, 2, 2] = 1
    src = np.stack([rng.uniform(0, 640, (npairs, n)), rng.uniform(0, 480, (npairs, n))], 2)
    p = np.einsum("pij,pnj->pni", H, np.concatenate([src, np.ones((npairs, n, 1))], 2))
    dst = p[..., :2] / p[..., 2:]
    inl = rng.random((npairs, n)) < inlier_ratio
    dst += inl[..., None] * rng.normal(0, 1.0, dst.shape)
    out = np.stack([rng.uniform(0, 640, (npairs, n)), rng.uniform(0, 480, (npairs, n))], 2)
    dst = np.where(inl[..., None], dst, out)
    offs = (np.arange(npairs + 1) * n).astype(np.int64)
    return src.reshape(-1, 2).astype(np.float32), dst.reshape(-1, 2).astype(np.float32), offs


# A frontal standing pose in OpenPose COCO order (pixel coordinates), used when no fixture poses are given.
CANONICAL_POSE = np.array([
    [320, 90], [320, 140], [275, 140], [262, 205], [255, 265], [365, 140], [378, 205], [385, 265],
    [292, 270], [288, 365], [285, 455], [348, 270], [352, 365], [355, 455], [310, 80], [330, 80],
    [297, 88], [343, 88]], dtype=np.float64)


def pose_database(base_poses, n, seed=3, undetected_fraction=0.1):
    """[n,18,2] float32: random base pose x random similarity (rot +-25 deg, scale 0.5..2, shift) + joint noise
    sigma in {0,2,5,15} px; `undetected_fraction` of the poses get 1..4 joints zeroed."""
    rng = np.random.default_rng(seed)
    base = np.asarray(base_poses, dtype=np.float64).reshape(-1, 18, 2)
    pick = rng.integers(0, len(base), n)
    P = base[pick].copy()
    und = (P[..., 0] == 0) & (P[..., 1] == 0)
    th = rng.uniform(-25, 25, n) * np.pi / 180
    sc = rng.uniform(0.5, 2.0, n)
    R = np.stack([np.stack([np.cos(th), -np.sin(th)], 1), np.stack([np.sin(th), np.cos(th)], 1)], 1) * sc[:, None, None]
    c = np.where(und[..., None], np.nan, P)
    centre = np.nanmean(c, axis=1, keepdims=True)
    P = np.einsum("nij,nkj->nki", R, P - centre) + centre + rng.uniform(-50, 50, (n, 1, 2)) + 400.0
    sigma = rng.choice([0.0, 2.0, 5.0, 15.0], n)
    P += rng.normal(0, 1, P.shape) * sigma[:, None, None]
    P = np.maximum(P, 1.0)  # keep detected joints off the (0,0) "undetected" sentinel
    P[und] = 0
    drop = np.nonzero(rng.random(n) < undetected_fraction)[0]
    for k in drop:
        P[k, rng.choice(18, rng.integers(1, 5), replace=False)] = 0
    return P.astype(np.float32)
