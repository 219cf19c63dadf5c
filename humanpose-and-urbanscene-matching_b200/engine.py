"""Batched entry points over torch CUDA tensors.

torch is the buffer carrier only: every function checks its arguments, allocates outputs and the
workspace as torch tensors on the inputs' device, and hands ``data_ptr()`` values plus the current CUDA
stream to the C ABI of libhpusm.so (include/hpusm.h).  There is no CPU path here: a non-CUDA tensor is
an error, and a failing library call raises ``HpusmError``.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

L2_AUTO, L2_SIMT, L2_TC_INT, L2_TC_SPLIT, L2_TC_I8, L2_TC_Q16 = (_lib.L2_AUTO, _lib.L2_SIMT, _lib.L2_TC_INT,
                                                                  _lib.L2_TC_SPLIT, _lib.L2_TC_I8, _lib.L2_TC_Q16)
POSE_DETAIL = _lib.POSE_DETAIL

# thresholds.py:7-13 of the reference, in the order hpusm_pose_match_batch expects
DEFAULT_POSE_THRESHOLDS = (0.236, 19.8, 0.082, 24.0, 0.125)


def _cuda(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _workspace(nbytes, dev):
    # torch's caching allocator hands out 512-byte aligned blocks; the ABI asks for 256
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def async_status():
    """HPUSM_OK (0) or the error code an asynchronous entry point recorded on the device since the last call (explicit
    L2_TC_I8 / L2_TC_INT on data outside their value range: every idx of that call is -1).  Call after synchronising."""
    return int(lib.hpusm_async_status())


def launch_count():
    """Kernels launched by libhpusm.so in this process so far."""
    return int(lib.hpusm_launch_count())


def profile_enable(on=True):
    """Bracket the dominant kernel of every path with CUDA events (roofline accounting in bench.py)."""
    check(lib.hpusm_profile_enable(1 if on else 0))


def profile_collect():
    """(summed milliseconds, launches) of the bracketed kernels since the last collect."""
    ms = ctypes.c_double(0.0)
    n = ctypes.c_int64(0)
    check(lib.hpusm_profile_collect(ctypes.cast(ctypes.byref(ms), ctypes.c_void_p),
                                    ctypes.cast(ctypes.byref(n), ctypes.c_void_p)))
    return ms.value, n.value


# ---- K2: Hamming ------------------------------------------------------------------------------------
def knn2_hamming(q, t):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) (features.py:59): (idx [nq,2] i32, dist [nq,2] i32)."""
    q = _cuda(q, torch.uint8, "q")
    t = _cuda(t, torch.uint8, "t")
    if q.dim() != 2 or t.dim() != 2 or q.shape[1] != t.shape[1]:
        raise ValueError("descriptor arrays must be [n, nbytes] with equal width")
    nq, nb = q.shape
    nt = t.shape[0]
    with torch.cuda.device(q.device):
        idx = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        dist = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        ws = _workspace(lib.hpusm_knn2_hamming_workspace_bytes(nq, nt, nb), q.device)
        check(lib.hpusm_knn2_hamming(_ptr(q), nq, _ptr(t), nt, nb, _ptr(idx), _ptr(dist), _ptr(ws), ws.numel(),
                                     _stream(q.device)))
    return idx, dist


HAMMING_AUTO, HAMMING_POPC, HAMMING_TC = 0, 1, 2
_HAMMING_ENGINES = {"auto": HAMMING_AUTO, "popc": HAMMING_POPC, "tc": HAMMING_TC}


def knn2_hamming_segmented(q, t, seg_offsets, ratio=0.8, want_best_idx=False, engine=HAMMING_AUTO):
    """Per database image: kNN-2 inside the image + Lowe ratio + survivor count (score [nseg] i32).
    engine: HAMMING_AUTO / HAMMING_POPC / HAMMING_TC (or 'auto' / 'popc' / 'tc'); same answer from both engines."""
    engine = _HAMMING_ENGINES.get(engine, engine)
    q = _cuda(q, torch.uint8, "q")
    t = _cuda(t, torch.uint8, "t")
    seg_offsets = _cuda(seg_offsets, torch.int64, "seg_offsets")
    nq, nb = q.shape
    nt = t.shape[0]
    nseg = seg_offsets.numel() - 1
    with torch.cuda.device(q.device):
        score = torch.empty((nseg,), dtype=torch.int32, device=q.device)
        best = torch.empty((nseg, nq), dtype=torch.int32, device=q.device) if want_best_idx else None
        ws = _workspace(lib.hpusm_knn2_hamming_segmented_workspace_bytes(nq, nt, nseg, nb, engine), q.device)
        check(lib.hpusm_knn2_hamming_segmented(_ptr(q), nq, _ptr(t), nt, _ptr(seg_offsets), nseg, nb, float(ratio), engine,
                                               _ptr(score), _ptr(best), _ptr(ws), ws.numel(), _stream(q.device)))
    return (score, best) if want_best_idx else score


# ---- K1: L2 ------------------------------------------------------------------------------------------
def knn2_l2(q, t, mode=L2_AUTO, out=None, workspace=None):
    """cv2.BFMatcher(NORM_L2).knnMatch(q, t, k=2) (features.py:59): (idx [nq,2] i32, dist [nq,2] f32 = sqrt)."""
    q = _cuda(q, torch.float32, "q")
    t = _cuda(t, torch.float32, "t")
    if q.dim() != 2 or t.dim() != 2 or q.shape[1] != t.shape[1]:
        raise ValueError("descriptor arrays must be [n, d] with equal d")
    nq, d = q.shape
    nt = t.shape[0]
    with torch.cuda.device(q.device):
        if out is None:
            idx = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
            dist = torch.empty((nq, 2), dtype=torch.float32, device=q.device)
        else:
            idx, dist = out
        need = lib.hpusm_knn2_l2_workspace_bytes(nq, nt, d, mode)
        ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, q.device)
        check(lib.hpusm_knn2_l2(_ptr(q), nq, _ptr(t), nt, d, mode, _ptr(idx), _ptr(dist), _ptr(ws), ws.numel(),
                                _stream(q.device)))
    return idx, dist


def knn2_l2_workspace(nq, nt, d, mode, device):
    return _workspace(lib.hpusm_knn2_l2_workspace_bytes(nq, nt, d, mode), device)


def knn2_l2_last_mode():
    return int(lib.hpusm_knn2_l2_last_mode())


def knn2_l2_last_fallback_rows():
    return int(lib.hpusm_knn2_l2_last_fallback_rows())


def knn2_merge(part_idx, part_dist):
    """Lexicographic (dist, idx) merge of [nparts, nq, 2] partial top-2 lists (indices already global)."""
    part_idx = _cuda(part_idx, torch.int32, "part_idx")
    nparts, nq, _ = part_idx.shape
    dev = part_idx.device
    with torch.cuda.device(dev):
        idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        if part_dist.dtype == torch.float32:
            part_dist = _cuda(part_dist, torch.float32, "part_dist")
            dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
            check(lib.hpusm_knn2_merge_f32(_ptr(part_idx), _ptr(part_dist), nparts, nq, _ptr(idx), _ptr(dist),
                                           _stream(dev)))
        else:
            part_dist = _cuda(part_dist, torch.int32, "part_dist")
            dist = torch.empty((nq, 2), dtype=torch.int32, device=dev)
            check(lib.hpusm_knn2_merge_i32(_ptr(part_idx), _ptr(part_dist), nparts, nq, _ptr(idx), _ptr(dist),
                                           _stream(dev)))
    return idx, dist


def ratio_filter(idx, dist, ratio=0.8):
    """features.filter_matches' test (features.py:48): (keep [nq] u8, count [1] i32 on device)."""
    idx = _cuda(idx, torch.int32, "idx")
    is_int = dist.dtype == torch.int32
    dist = _cuda(dist, torch.int32 if is_int else torch.float32, "dist")
    nq = idx.shape[0]
    dev = idx.device
    with torch.cuda.device(dev):
        keep = torch.empty((nq,), dtype=torch.uint8, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        check(lib.hpusm_ratio_filter(_ptr(dist), _ptr(idx), nq, 1 if is_int else 0, float(ratio), _ptr(keep),
                                     _ptr(count), _stream(dev)))
    return keep, count


def gather_matches(kp_q, kp_t, idx, keep):
    """Order-preserving compaction + keypoint gather (features.py:49-53): (p_q, p_t, sel, count)."""
    kp_q = _cuda(kp_q, torch.float32, "kp_q")
    kp_t = _cuda(kp_t, torch.float32, "kp_t")
    idx = _cuda(idx, torch.int32, "idx")
    keep = _cuda(keep, torch.uint8, "keep")
    nq = idx.shape[0]
    dev = idx.device
    with torch.cuda.device(dev):
        p_q = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        p_t = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        sel = torch.empty((nq,), dtype=torch.int32, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        ws = _workspace(lib.hpusm_gather_matches_workspace_bytes(nq), dev)
        check(lib.hpusm_gather_matches(_ptr(kp_q), _ptr(kp_t), _ptr(idx), _ptr(keep), nq, _ptr(p_q), _ptr(p_t),
                                       _ptr(sel), _ptr(count), _ptr(ws), ws.numel(), _stream(dev)))
    return p_q, p_t, sel, count


def segment_matches(best_idx, count, kp_q, kp_t, min_count=16):
    """Per-image match lists after knn2_hamming_segmented(..., want_best_idx=True): (pair_offsets [nseg+1] i64,
    src [total,2], dst [total,2], sel [total] i32); images below min_count (features.py:64) contribute nothing."""
    best_idx = _cuda(best_idx, torch.int32, "best_idx")
    count = _cuda(count, torch.int32, "count")
    kp_q = _cuda(kp_q, torch.float32, "kp_q")
    kp_t = _cuda(kp_t, torch.float32, "kp_t")
    nseg, nq = best_idx.shape
    dev = best_idx.device
    with torch.cuda.device(dev):
        offs = torch.empty((nseg + 1,), dtype=torch.int64, device=dev)
        cap = int(count[count >= int(min_count)].sum().item())  # sizes the outputs (the only host read of the stage)
        room = max(cap, 1)  # no image above min_count: the ABI still wants non-null outputs
        src = torch.empty((room, 2), dtype=torch.float32, device=dev)
        dst = torch.empty((room, 2), dtype=torch.float32, device=dev)
        sel = torch.empty((room,), dtype=torch.int32, device=dev)
        check(lib.hpusm_segment_matches(_ptr(best_idx), _ptr(count), nseg, nq, _ptr(kp_q), _ptr(kp_t), int(min_count),
                                        _ptr(offs), _ptr(src), _ptr(dst), _ptr(sel), _stream(dev)))
    return offs, src[:cap], dst[:cap], sel[:cap]


def validate_homography_batch(H):
    """features.validate_homography (features.py:177-211) on [n,3,3] float64: valid [n] u8."""
    H = _cuda(H, torch.float64, "H")
    n = H.shape[0]
    with torch.cuda.device(H.device):
        valid = torch.empty((n,), dtype=torch.uint8, device=H.device)
        check(lib.hpusm_validate_homography_batch(_ptr(H), n, _ptr(valid), _stream(H.device)))
    return valid


# ---- K3: RANSAC ----------------------------------------------------------------------------------------
SAMPLER_COUNTER = 0  # counter-based sample table keyed on (seed, pair, iteration, attempt): parallel draws
SAMPLER_CV2 = 1      # OpenCV's own fixed-seed sample sequence: the result is cv2.findHomography's


def _sampler(sampler):
    if sampler in (SAMPLER_COUNTER, "counter"):
        return SAMPLER_COUNTER
    if sampler in (SAMPLER_CV2, "cv2"):
        return SAMPLER_CV2
    raise ValueError("sampler must be 'counter' or 'cv2', got %r" % (sampler,))


def ransac_homography_batch(src, dst, pair_offsets, thresh=5.0, nhyp=2000, confidence=0.995, seed=0, refine=True,
                            pair_index0=0, sampler=SAMPLER_COUNTER):
    """cv2.findHomography(src, dst, RANSAC, thresh) per pair (features.py:66,:69).

    Returns (H [npairs,3,3] f64, mask [total] u8, ninl [npairs] i32, best_hyp [npairs] i32).  `sampler`: 'cv2' replays
    OpenCV's own sample sequence (mask and H are cv2.findHomography's), 'counter' draws from the seeded counter-based
    table; `pair_index0`: the index of this call's first pair in that table (a slice of a larger batch draws the larger
    batch's samples)."""
    src = _cuda(src, torch.float32, "src")
    dst = _cuda(dst, torch.float32, "dst")
    pair_offsets = _cuda(pair_offsets, torch.int64, "pair_offsets")
    total = src.shape[0]
    npairs = pair_offsets.numel() - 1
    dev = src.device
    sampler = _sampler(sampler)
    with torch.cuda.device(dev):
        H = torch.empty((npairs, 3, 3), dtype=torch.float64, device=dev)
        mask = torch.empty((total,), dtype=torch.uint8, device=dev)
        ninl = torch.empty((npairs,), dtype=torch.int32, device=dev)
        best = torch.empty((npairs,), dtype=torch.int32, device=dev)
        ws = _workspace(lib.hpusm_ransac_homography_workspace_bytes(total, npairs, nhyp, sampler), dev)
        check(lib.hpusm_ransac_homography_batch(_ptr(src), _ptr(dst), _ptr(pair_offsets), npairs, total, float(thresh),
                                                int(nhyp), float(confidence), sampler, int(seed), int(pair_index0),
                                                1 if refine else 0, _ptr(H), _ptr(mask), _ptr(ninl), _ptr(best),
                                                _ptr(ws), ws.numel(), _stream(dev)))
    return H, mask, ninl, best


def ransac_score_batch(src, dst, pair_offsets, thresh=5.0, nhyp=1024, seed=0, out=None, pair_index0=0,
                       sampler=SAMPLER_COUNTER):
    """Inlier count of every iteration's model of every pair: counts [npairs, nhyp] i32 (-1 = no acceptable sample)."""
    src = _cuda(src, torch.float32, "src")
    dst = _cuda(dst, torch.float32, "dst")
    pair_offsets = _cuda(pair_offsets, torch.int64, "pair_offsets")
    total = src.shape[0]
    npairs = pair_offsets.numel() - 1
    dev = src.device
    sampler = _sampler(sampler)
    with torch.cuda.device(dev):
        counts = out if out is not None else torch.empty((npairs, nhyp), dtype=torch.int32, device=dev)
        ws = _workspace(lib.hpusm_ransac_score_workspace_bytes(total, npairs, nhyp, sampler), dev)
        check(lib.hpusm_ransac_score_batch(_ptr(src), _ptr(dst), _ptr(pair_offsets), npairs, total, float(thresh),
                                           int(nhyp), sampler, int(seed), int(pair_index0), _ptr(counts), _ptr(ws),
                                           ws.numel(), _stream(dev)))
    return counts


# ---- ORB description (the step in front of K2) --------------------------------------------------------------
ORB_LEVELS = 8                           # features.py:19-20
ORB_SCALE = float(np.float32(1.2))       # ORB::create takes a float: OpenCV works with the double 1.2000000476837158
ORB_FAST_THRESHOLD, ORB_EDGE, ORB_PATCH, ORB_NFEATURES = 20, 31, 31, 100000


def orb_keypoint_params(pts, octaves, angles, scale_factor=ORB_SCALE):
    """Host-side per-keypoint inputs of hpusm_orb_compute from cv2.KeyPoint fields: (kp_lxy [n,3] int32, kp_cs [n,2] f32).
    float32 arithmetic as in OpenCV's computeOrbDescriptors: scale = 1.f / (float)pow(scaleFactor, octave), centre =
    cvRound(pt * scale), (cos, sin) of angle * (float)(pi/180)."""
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    octaves = np.asarray(octaves, np.int32).reshape(-1)
    f32 = np.float32
    level_scale = np.power(np.float64(scale_factor), octaves.astype(np.float64)).astype(np.float32)
    inv = f32(1.0) / level_scale
    lxy = np.empty((len(pts), 3), np.int32)
    lxy[:, 0] = octaves
    lxy[:, 1] = np.rint(pts[:, 0] * inv).astype(np.int32)
    lxy[:, 2] = np.rint(pts[:, 1] * inv).astype(np.int32)
    ang = (np.asarray(angles, np.float32).reshape(-1) * f32(np.pi / 180.0)).astype(np.float64)
    cs = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)
    return lxy, cs


def orb_compute(img, kp_lxy, kp_cs, nlevels=ORB_LEVELS, scale_factor=ORB_SCALE, workspace=None):
    """cv2.ORB.compute on the device: img [h,w] uint8 CUDA tensor, keypoint parameters from orb_keypoint_params (CUDA
    tensors).  Returns desc [n,32] uint8 on the device (bit-identical to cv2's), ready for knn2_hamming."""
    img = _cuda(img, torch.uint8, "img")
    kp_lxy = _cuda(kp_lxy, torch.int32, "kp_lxy")
    kp_cs = _cuda(kp_cs, torch.float32, "kp_cs")
    if img.dim() != 2:
        raise ValueError("img must be [height, width] uint8")
    h, w = img.shape
    n = kp_lxy.shape[0]
    dev = img.device
    with torch.cuda.device(dev):
        desc = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        need = lib.hpusm_orb_workspace_bytes(w, h, int(nlevels), float(scale_factor))
        ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, dev)
        check(lib.hpusm_orb_compute(_ptr(img), w, h, w, int(nlevels), float(scale_factor), _ptr(kp_lxy), _ptr(kp_cs), n,
                                    _ptr(desc), _ptr(ws), ws.numel(), _stream(dev)))
    return desc


def orb_level_quotas(nlevels=ORB_LEVELS, nfeatures=ORB_NFEATURES, scale_factor=ORB_SCALE):
    """nfeaturesPerLevel of OpenCV's computeKeyPoints (float32 arithmetic): the per-level caps on the keypoint count."""
    f32 = np.float32
    factor = f32(1.0 / scale_factor)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(np.power(np.float64(factor), np.float64(nlevels))))
    out, total = [], 0
    for _ in range(nlevels - 1):
        out.append(int(np.rint(nd)))
        total += out[-1]
        nd = f32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def orb_detect_and_compute(img, capacity=None, want_descriptors=True, nlevels=ORB_LEVELS, scale_factor=ORB_SCALE):
    """cv2.ORB.detectAndCompute(img, None) of the reference's ORB on the device.  img [h,w] uint8 CUDA tensor.
    Returns a dict of device tensors cut to the keypoint count: pt [n,2] f32, octave [n] i32, angle / response / size
    [n] f32, desc [n,32] u8 (None when not wanted), plus level_counts (host list), and over_quota (a level exceeded
    OpenCV's per-level cap: cv2 would have culled by response with an unspecified tie order, the device path does not -- the
    drop-in raises unless the caller opted into OpenCV's detector)."""
    img = _cuda(img, torch.uint8, "img")
    if img.dim() != 2:
        raise ValueError("img must be [height, width] uint8")
    h, w = img.shape
    dev = img.device
    cap = int(capacity) if capacity else max(4096, (h * w) // 8)
    with torch.cuda.device(dev):
        f32, i32 = torch.float32, torch.int32
        xy = torch.empty((cap, 2), dtype=f32, device=dev)
        octave = torch.empty((cap,), dtype=i32, device=dev)
        angle = torch.empty((cap,), dtype=f32, device=dev)
        response = torch.empty((cap,), dtype=f32, device=dev)
        size = torch.empty((cap,), dtype=f32, device=dev)
        lxy = torch.empty((cap, 3), dtype=i32, device=dev)
        cs = torch.empty((cap, 2), dtype=f32, device=dev)
        desc = torch.empty((cap, 32), dtype=torch.uint8, device=dev) if want_descriptors else None
        counts = torch.empty((nlevels + 1,), dtype=i32, device=dev)
        ws = _workspace(lib.hpusm_orb_detect_workspace_bytes(w, h, int(nlevels), float(scale_factor)), dev)
        check(lib.hpusm_orb_detect_and_compute(_ptr(img), w, h, w, int(nlevels), float(scale_factor), ORB_FAST_THRESHOLD,
                                               ORB_EDGE, ORB_PATCH, cap, _ptr(xy), _ptr(octave), _ptr(angle), _ptr(response),
                                               _ptr(size), _ptr(lxy), _ptr(cs), _ptr(desc), _ptr(counts), _ptr(ws),
                                               ws.numel(), _stream(dev)))
        level_counts = counts.cpu().tolist()  # the one host read: how many keypoints there are
    n = level_counts[-1]
    if n > cap:
        return orb_detect_and_compute(img, capacity=n, want_descriptors=want_descriptors, nlevels=nlevels,
                                      scale_factor=scale_factor)
    quotas = orb_level_quotas(nlevels, ORB_NFEATURES, scale_factor)
    return {"pt": xy[:n], "octave": octave[:n], "angle": angle[:n], "response": response[:n], "size": size[:n],
            "kp_lxy": lxy[:n], "kp_cs": cs[:n], "desc": desc[:n] if desc is not None else None,
            "level_counts": level_counts[:-1], "over_quota": any(c > q for c, q in zip(level_counts[:-1], quotas))}


def perspective_transform(pts, H):
    """cv2.perspectiveTransform(pts, H) (features.py:77, transformation.py:38)."""
    pts = _cuda(pts, torch.float32, "pts")
    H = _cuda(H, torch.float64, "H")
    n = pts.numel() // 2
    dev = pts.device
    with torch.cuda.device(dev):
        out = torch.empty_like(pts)
        check(lib.hpusm_perspective_transform(_ptr(pts), n, _ptr(H), _ptr(out), _stream(dev)))
    return out


# ---- K4: pose ------------------------------------------------------------------------------------------
def affine_lsq_batch(model, inp):
    """common.find_transformation batched (common.py:37-101): (transform [n,npts,2], A [n,3,3], has_A [n] u8)."""
    model = _cuda(model, torch.float64, "model")
    inp = _cuda(inp, torch.float64, "input")
    n, npts, _ = inp.shape
    dev = inp.device
    with torch.cuda.device(dev):
        out = torch.empty_like(inp)
        A = torch.empty((n, 3, 3), dtype=torch.float64, device=dev)
        has = torch.empty((n,), dtype=torch.uint8, device=dev)
        check(lib.hpusm_affine_lsq_batch(_ptr(model), _ptr(inp), n, npts, _ptr(out), _ptr(A), _ptr(has), _stream(dev)))
    return out, A, has


def procrustes_batch(X, Y, scaling=True, reflection="best"):
    """procrustes.procrustes batched (procrustes.py:150-259): (d [n], Z [n,npts,2], tform [n,7])."""
    X = _cuda(X, torch.float64, "X")
    Y = _cuda(Y, torch.float64, "Y")
    n, npts, _ = X.shape
    refl = 0 if (isinstance(reflection, str) and reflection == "best") else (1 if reflection else 2)
    dev = X.device
    with torch.cuda.device(dev):
        d = torch.empty((n,), dtype=torch.float64, device=dev)
        Z = torch.empty_like(X)
        tform = torch.empty((n, 7), dtype=torch.float64, device=dev)
        check(lib.hpusm_procrustes_batch(_ptr(X), _ptr(Y), n, npts, 1 if scaling else 0, refl, _ptr(d), _ptr(Z),
                                         _ptr(tform), _stream(dev)))
    return d, Z, tform


def pose_match_batch(query, db, query_is_model=True, thresholds=DEFAULT_POSE_THRESHOLDS, want_detail=False, out=None,
                     exact=False, workspace=None):
    """single_person.match_single of one query against a pose database (single_person.py:80-189).

    float32 databases are swept by the FP32 fast path with the exact FP64 kernel behind it for every pose it cannot
    certify (same decisions as the FP64 kernel, scores within 2e-6); `exact=True` runs the FP64 kernel on every pose.
    Returns (match [n] u8, score [n] f32[, detail [n, POSE_DETAIL] f64])."""
    f64 = db.dtype == torch.float64
    query = _cuda(query, torch.float64 if f64 else torch.float32, "query")
    db = _cuda(db, torch.float64 if f64 else torch.float32, "db")
    if query.numel() != 36 or db.dim() != 3 or db.shape[1:] != (18, 2):
        raise ValueError("query must be [18,2] and db [n,18,2]")
    n = db.shape[0]
    dev = db.device
    thr = (ctypes.c_double * 5)(*[float(x) for x in thresholds])
    with torch.cuda.device(dev):
        if out is None:
            match = torch.empty((n,), dtype=torch.uint8, device=dev)
            score = torch.empty((n,), dtype=torch.float32, device=dev)
        else:
            match, score = out
        detail = torch.empty((n, POSE_DETAIL), dtype=torch.float64, device=dev) if want_detail else None
        if f64:
            check(lib.hpusm_pose_match_batch_f64(_ptr(query), _ptr(db), n, 1 if query_is_model else 0,
                                                 ctypes.cast(thr, ctypes.c_void_p), _ptr(match), _ptr(score), _ptr(detail),
                                                 _stream(dev)))
        else:
            ws = None
            if not exact and not want_detail:
                need = lib.hpusm_pose_match_workspace_bytes(n)
                ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, dev)
            check(lib.hpusm_pose_match_batch(_ptr(query), _ptr(db), n, 1 if query_is_model else 0,
                                             ctypes.cast(thr, ctypes.c_void_p), _ptr(match), _ptr(score), _ptr(detail),
                                             _ptr(ws), ws.numel() if ws is not None else 0, _stream(dev)))
    return (match, score, detail) if want_detail else (match, score)


def pose_match_pairs(model, inp, thresholds=DEFAULT_POSE_THRESHOLDS, want_detail=False):
    """match_single(model[i], input[i]) for n arbitrary pairs in one launch (multi_person.py:169-207 candidate stage).
    model, inp [n,18,2] float32 or float64.  Returns (match [n] u8, score [n] f32[, detail [n, POSE_DETAIL] f64])."""
    f64 = inp.dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    model = _cuda(model, dt, "model")
    inp = _cuda(inp, dt, "input")
    if model.shape != inp.shape or inp.dim() != 3 or inp.shape[1:] != (18, 2):
        raise ValueError("model and input must both be [n,18,2]")
    n = inp.shape[0]
    dev = inp.device
    thr = (ctypes.c_double * 5)(*[float(x) for x in thresholds])
    with torch.cuda.device(dev):
        match = torch.empty((n,), dtype=torch.uint8, device=dev)
        score = torch.empty((n,), dtype=torch.float32, device=dev)
        detail = torch.empty((n, POSE_DETAIL), dtype=torch.float64, device=dev) if want_detail else None
        check(lib.hpusm_pose_match_pairs(_ptr(model), _ptr(inp), n, 1 if f64 else 0, ctypes.cast(thr, ctypes.c_void_p),
                                         _ptr(match), _ptr(score), _ptr(detail), _stream(dev)))
    return (match, score, detail) if want_detail else (match, score)


def scene_score_batch(src, dst, pair_offsets, mask, H2, model_pose, input_pose, model_w, model_h, seed=0, picks=None):
    """Perspective correction + affine_multi score per image pair (urban_scene.py:80-145).  model_pose [18,2] (shared)
    or [npairs,18,2]; input_pose [npairs,18,2]; float64.  Returns (score [npairs] f64, picks [npairs,5] i32)."""
    src = _cuda(src, torch.float32, "src")
    dst = _cuda(dst, torch.float32, "dst")
    pair_offsets = _cuda(pair_offsets, torch.int64, "pair_offsets")
    mask = _cuda(mask, torch.uint8, "mask")
    H2 = _cuda(H2, torch.float64, "H2")
    model_pose = _cuda(model_pose, torch.float64, "model_pose")
    input_pose = _cuda(input_pose, torch.float64, "input_pose")
    npairs = pair_offsets.numel() - 1
    nj = input_pose.shape[-2]
    shared = model_pose.dim() == 2
    dev = src.device
    if picks is not None:
        picks = _cuda(picks, torch.int32, "picks")
    with torch.cuda.device(dev):
        score = torch.empty((npairs,), dtype=torch.float64, device=dev)
        used = torch.full((npairs, 5), -1, dtype=torch.int32, device=dev)
        check(lib.hpusm_scene_score_batch(_ptr(src), _ptr(dst), _ptr(pair_offsets), _ptr(mask), _ptr(H2), _ptr(model_pose),
                                          1 if shared else 0, _ptr(input_pose), npairs, nj, float(model_w), float(model_h),
                                          int(seed), _ptr(picks), _ptr(score), _ptr(used), _stream(dev)))
    return score, used


def feature_scaling_batch(pts):
    """common.feature_scaling on [n, npts, 2] float64 point sets (common.py:688-708)."""
    pts = _cuda(pts, torch.float64, "pts")
    n, npts, _ = pts.shape
    with torch.cuda.device(pts.device):
        out = torch.empty_like(pts)
        check(lib.hpusm_feature_scaling_batch(_ptr(pts), n, npts, _ptr(out), _stream(pts.device)))
    return out


def pose_topk(match, score, k=16, index_offset=0):
    """k best-scoring matching poses: (top_score [k] f32, top_index [k] i64, match_count [1] i32)."""
    match = _cuda(match, torch.uint8, "match")
    score = _cuda(score, torch.float32, "score")
    n = match.numel()
    dev = match.device
    with torch.cuda.device(dev):
        ts = torch.empty((k,), dtype=torch.float32, device=dev)
        ti = torch.empty((k,), dtype=torch.int64, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        ws = _workspace(lib.hpusm_pose_topk_workspace_bytes(n, k), dev)
        check(lib.hpusm_pose_topk(_ptr(match), _ptr(score), n, int(index_offset), int(k), _ptr(ts), _ptr(ti), _ptr(cnt),
                                  _ptr(ws), ws.numel(), _stream(dev)))
    return ts, ti, cnt


def pose_topk_merge(cand_score, cand_index, k=16):
    """k lexicographically smallest (score, index) of up to 4096 candidates (index < 0 = padding): (score [k], index [k])."""
    cand_score = _cuda(cand_score.reshape(-1), torch.float32, "cand_score")
    cand_index = _cuda(cand_index.reshape(-1), torch.int64, "cand_index")
    dev = cand_score.device
    with torch.cuda.device(dev):
        ts = torch.empty((k,), dtype=torch.float32, device=dev)
        ti = torch.empty((k,), dtype=torch.int64, device=dev)
        check(lib.hpusm_pose_topk_merge(_ptr(cand_score), _ptr(cand_index), cand_score.numel(), int(k), _ptr(ts), _ptr(ti),
                                        _stream(dev)))
    return ts, ti


def microbench_mma(kind, rounds, device):
    """Back-to-back tcgen05.mma on every SM (kind "bf16" | "i8" at N = 256, both operands in shared memory; "i8n64",
    "i8n64t", "i8n128", "i8n128t": u8 at the kNN engines' shapes, t = A operand in tensor memory); returns the
    operations issued (2 x MACs) for timing."""
    k, n = {"bf16": (0, 256), "i8": (1, 256), "i8n64": (2, 64), "i8n64t": (3, 64), "i8n128": (4, 128), "i8n128t": (5, 128)}[kind]
    with torch.cuda.device(device):
        check(lib.hpusm_microbench_mma(k, int(rounds), _stream(device)))
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    return 2.0 * 128 * n * (16 if k == 0 else 32) * 64 * int(rounds) * sms


def microbench_popc(blocks, threads, iters, device, op="popc"):
    """op: "popc" | "lop3" | "imax3" | "fmax3" | "imax" | "fmax" | "ffma" | "ffma2" -- every thread issues `iters` of
    that instruction."""
    sink = torch.zeros((1,), dtype=torch.int32, device=device)
    if op in ("ffma", "ffma2"):
        with torch.cuda.device(device):
            check(lib.hpusm_microbench_ffma(1 if op == "ffma2" else 0, int(blocks), int(threads), int(iters), _ptr(sink),
                                            _stream(device)))
        return
    minmax = {"imax3": 0, "fmax3": 1, "imax": 2, "fmax": 3}
    with torch.cuda.device(device):
        if op in minmax:
            check(lib.hpusm_microbench_minmax(minmax[op], int(blocks), int(threads), int(iters), _ptr(sink), _stream(device)))
            return
        fn = lib.hpusm_microbench_popc if op == "popc" else lib.hpusm_microbench_lop3
        check(fn(int(blocks), int(threads), int(iters), _ptr(sink), _stream(device)))
