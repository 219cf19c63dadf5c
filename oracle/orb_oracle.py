"""CPU restatement of cv2.ORB.compute for given keypoints -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

The reference describes every image with `detector.detectAndCompute` (urbanscene/urban_scene.py:31-32), the detector being
cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
scoreType=ORB_FAST_SCORE, patchSize=31) (urbanscene/features.py:19-20).  The arithmetic lives in OpenCV
(modules/features2d/src/orb.cpp, imgproc resize / filter; NOT vendored under /root/reference; binary here:
opencv-python-headless 4.13.0.92).  The description half (ORB_Impl::detectAndCompute with given keypoints ->
computeOrbDescriptors) is restated from the published algorithm and pinned on cv2's own output:
  * pyramid: level L has size (cvRound(cols / s_L), cvRound(rows / s_L)), s_L = (float)pow(1.2, L), and is
    resize(level L-1, INTER_LINEAR_EXACT): separable bilinear with Q8.8 coefficients (fraction rounded to 1/256), 16.16
    accumulation, one final rounding -- bit-identical to cv2.resize on every level of every fixture;
  * every level is smoothed by GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101).  Inside ORB OpenCV does NOT take its
    fixed-point "bit-exact" Gaussian: the level is a view into the pyramid atlas and goes through the generic separable
    filter, uint8 -> float32 rows -> uint8, with the float32 kernel of getGaussianKernel and fused multiply-adds: rows as
    a left-to-right FMA chain, columns as the symmetric form k3*c + k2*(u1+d1) + k1*(u2+d2) + k0*(u3+d3), cvRound at the
    end (probe: 0 differing pixels in 17.9 M against cv2.sepFilter2D, which ORB's output pins through the descriptors);
  * descriptor: for keypoint (pt, octave, angle), scale = 1.f / s_octave, centre (cvRound(pt.x * scale),
    cvRound(pt.y * scale)), a = cos(angle * pi/180), b = sin(..) in float32; test k compares the smoothed level at
    (cvRound(x*a - y*b), cvRound(x*b + y*a)) for the two points of pair k of the 256-pair pattern; bit = t0 < t1 (WTA_K = 2).
The 256 x 2 test points (OpenCV's learned bit_pattern_31_) were read off the binary itself: a keypoint at angle 0 on a
black image with ONE pixel of value 11 (which the 7x7 Gaussian turns into a single pixel of value 1) sets exactly the
bits whose second point is that pixel; the inverted image gives the first points (oracle/record_golden.py
record_orb_pattern; stored in tests/golden/orb_pattern.npy).
Pinned by tests/test_oracle_golden.py::test_orb_* : every descriptor of the recorded fixtures is bit-identical to cv2's.
"""
import os

import numpy as np

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN_PATH = os.path.join(os.path.dirname(_HERE), "tests", "golden", "orb_pattern.npy")
# cv2.getGaussianKernel(7, 2, CV_32F): k[0..3] (symmetric), as float32 bit patterns
GAUSS7 = np.array([0x3D8FAFB1, 0x3E06387E, 0x3E434A39, 0x3E5D4AE0], np.uint32).view(np.float32)
NLEVELS = 8
SCALE_FACTOR = float(np.float32(1.2))  # ORB::create takes a float; ORB_Impl keeps it as the double 1.2000000476837158
FAST_THRESHOLD, EDGE_THRESHOLD, PATCH_SIZE, NFEATURES = 20, 31, 31, 100000  # features.py:19-20 (fastThreshold: cv2 default)


def level_scale(level):
    """getScale(): (float)pow(scaleFactor, level - firstLevel), firstLevel = 0."""
    return F32(np.power(np.float64(SCALE_FACTOR), np.float64(level)))


def level_size(w, h, level):
    s = level_scale(level)
    return int(np.rint(F32(w) / s)), int(np.rint(F32(h) / s))


def _coeffs(src_n, dst_n):
    """INTER_LINEAR_EXACT source index and Q8.8 weight of the right neighbour for every destination index."""
    scale = src_n / dst_n
    d = np.arange(dst_n, dtype=np.float64)
    f = (d + 0.5) * scale - 0.5
    i = np.floor(f).astype(np.int64)
    a = f - i
    lo = i < 0
    hi = i >= src_n - 1
    i = np.where(lo, 0, np.where(hi, src_n - 1, i))
    a = np.where(lo | hi, 0.0, a)
    return i, np.floor(a * 256 + 0.5).astype(np.int64)


def resize_linear_exact(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for uint8 single-channel images."""
    h, w = src.shape
    ix, ax = _coeffs(w, dw)
    iy, ay = _coeffs(h, dh)
    s = src.astype(np.int64)
    x1 = np.minimum(ix + 1, w - 1)
    rows = (256 - ax)[None, :] * s[:, ix] + ax[None, :] * s[:, x1]        # Q8.8
    y1 = np.minimum(iy + 1, h - 1)
    v = (256 - ay)[:, None] * rows[iy, :] + ay[:, None] * rows[y1, :]     # Q16.16
    return ((v + 32768) >> 16).astype(np.uint8)


def _fma(a, b, c):
    """fused multiply-add on float32 arrays: the product of two float32 is exact in float64."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def gaussian_blur7(img):
    """GaussianBlur(img, (7,7), 2, 2, BORDER_REFLECT_101) as ORB applies it (generic separable filter, float32, FMA)."""
    h, w = img.shape
    p = np.pad(img, 3, mode="reflect").astype(np.float32)
    k0, k1, k2, k3 = GAUSS7
    taps = [k0, k1, k2, k3, k2, k1, k0]
    r = p[:, 0:w] * taps[0]
    for i in range(1, 7):
        r = _fma(p[:, i:i + w], taps[i], r)                               # rows: left-to-right FMA chain
    c = r[3:3 + h] * k3
    c = _fma(r[2:2 + h] + r[4:4 + h], k2, c)                              # columns: symmetric pairs, centre first
    c = _fma(r[1:1 + h] + r[5:5 + h], k1, c)
    c = _fma(r[0:h] + r[6:6 + h], k0, c)
    return np.clip(np.rint(c), 0, 255).astype(np.uint8)


def pyramid(img, nlevels=NLEVELS):
    """Smoothed pyramid levels ORB samples its descriptors from."""
    h, w = img.shape
    raw = [np.ascontiguousarray(img, np.uint8)]
    for level in range(1, nlevels):
        dw, dh = level_size(w, h, level)
        raw.append(resize_linear_exact(raw[-1], dw, dh))
    return [gaussian_blur7(r) for r in raw]


def keypoint_params(pts, octaves, angles):
    """What the descriptor kernel consumes per keypoint: centre (cx, cy) in its level, cos and sin of its angle (float32)."""
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    octaves = np.asarray(octaves, np.int32)
    scale = F32(1.0) / np.array([level_scale(int(o)) for o in octaves], np.float32)
    cx = np.rint(pts[:, 0] * scale).astype(np.int32)
    cy = np.rint(pts[:, 1] * scale).astype(np.int32)
    ang = np.asarray(angles, np.float32) * F32(np.pi / 180.0)
    a = np.cos(ang.astype(np.float64)).astype(np.float32)
    b = np.sin(ang.astype(np.float64)).astype(np.float32)
    return cx, cy, a, b


# ---- detection: FAST-9/16 with non-maximum suppression per level, border filter, intensity-centroid angle ---------------
# Restated from ORB_Impl::detectAndCompute -> computeKeyPoints (features2d/src/orb.cpp), FAST_t<16> / cornerScore<16>
# (features2d/src/fast.cpp, fast_score.cpp), ICAngles (orb.cpp) and cv::fastAtan2 (core/src/mathfuncs_core): pinned on
# cv2.ORB.detect itself -- same keypoints in the same order with identical pt, octave, size, response and angle
# (tests/test_oracle_golden.py::test_orb_oracle_detect_*).
FAST_CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
               (-3, 0), (-3, 1), (-2, 2), (-1, 3)]  # (dx, dy), one way round the radius-3 circle


def fast_scores(img, threshold=FAST_THRESHOLD):
    """cornerScore<16> where the pixel is a FAST-9 corner, else 0.  A pixel (3 px inside the border) is a corner when 9
    contiguous circle pixels are all darker than v - t or all brighter than v + t; its score is the largest t' for which it
    still is, minus 1: max over the 16 arcs of min |v - c| on the arc (one polarity), minus 1."""
    h, w = img.shape
    out = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return out
    v = img[3:h - 3, 3:w - 3].astype(np.int32)
    d = np.stack([v - img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int32) for dx, dy in FAST_CIRCLE])
    d = np.concatenate([d, d[:8]], 0)
    s_dark = np.full(v.shape, -1000, np.int32)
    s_bright = np.full(v.shape, -1000, np.int32)
    for k in range(16):
        arc = d[k:k + 9]
        s_dark = np.maximum(s_dark, arc.min(0))
        s_bright = np.maximum(s_bright, (-arc).min(0))
    s = np.maximum(s_dark, s_bright)
    out[3:h - 3, 3:w - 3] = np.where(s > threshold, s - 1, 0)
    return out


def nonmax(score):
    """A corner survives when its score is strictly greater than the scores of its 8 neighbours (0 for a non-corner)."""
    h, w = score.shape
    p = np.pad(score, 1)
    keep = score > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= score > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    return keep


def umax_table(half_patch=PATCH_SIZE // 2):
    """Half-widths of the rows of ORB's circular patch (the symmetric construction of computeKeyPoints)."""
    um = np.zeros(half_patch + 2, np.int64)
    vmax = int(np.floor(half_patch * np.sqrt(F32(2.0)) / 2 + 1))
    vmin = int(np.ceil(half_patch * np.sqrt(F32(2.0)) / 2))
    for v in range(vmax + 1):
        um[v] = int(np.rint(np.sqrt(float(half_patch * half_patch - v * v))))
    v0 = 0
    for v in range(half_patch, vmin - 1, -1):
        while um[v0] == um[v0 + 1]:
            v0 += 1
        um[v] = v0
        v0 += 1
    return um[:half_patch + 1]


_K = F32(180 / np.pi)
_P1, _P3 = F32(0.9997878412794807) * _K, F32(-0.3258083974640975) * _K
_P5, _P7 = F32(0.1555786518463281) * _K, F32(-0.04432655554792128) * _K
_EPS = F32(np.finfo(np.float64).eps)


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees), float32 arithmetic, one rounding per operation; elementwise on arrays."""
    y = np.asarray(y, np.float32)
    x = np.asarray(x, np.float32)
    ax, ay = np.abs(x), np.abs(y)
    swap = ax < ay
    num = np.where(swap, ax, ay)
    den = np.where(swap, ay, ax) + _EPS
    c = (num / den).astype(np.float32)
    c2 = (c * c).astype(np.float32)
    a = ((((_P7 * c2 + _P5).astype(np.float32) * c2 + _P3).astype(np.float32) * c2 + _P1).astype(np.float32) * c).astype(np.float32)
    a = np.where(swap, F32(90.0) - a, a).astype(np.float32)
    a = np.where(x < 0, F32(180.0) - a, a).astype(np.float32)
    a = np.where(y < 0, F32(360.0) - a, a).astype(np.float32)
    return a


def ic_angles(img, xs, ys):
    """ICAngles: orientation from the intensity centroid of the circular patch around (x, y) of the UNSMOOTHED level."""
    um = umax_table()
    hp = PATCH_SIZE // 2
    im = img.astype(np.int64)
    m01 = np.zeros(len(xs), np.int64)
    m10 = np.zeros(len(xs), np.int64)
    for u in range(-hp, hp + 1):
        m10 += u * im[ys, xs + u]
    for v in range(1, hp + 1):
        d = int(um[v])
        vs = np.zeros(len(xs), np.int64)
        for u in range(-d, d + 1):
            plus, minus = im[ys + v, xs + u], im[ys - v, xs + u]
            vs += plus - minus
            m10 += u * (plus + minus)
        m01 += v * vs
    return fast_atan2(m01.astype(np.float32), m10.astype(np.float32))


def level_quotas(nlevels=NLEVELS, nfeatures=NFEATURES):
    """nfeaturesPerLevel of computeKeyPoints (float32 arithmetic)."""
    factor = F32(1.0 / SCALE_FACTOR)
    nd = F32(nfeatures) * (F32(1) - factor) / (F32(1) - F32(np.power(np.float64(factor), np.float64(nlevels))))
    out, total = [], 0
    for _ in range(nlevels - 1):
        out.append(int(np.rint(nd)))
        total += out[-1]
        nd = F32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def detect(img, nlevels=NLEVELS):
    """cv2.ORB.detect(img, None) for the reference's ORB: dict of arrays pt [n,2] f32, octave [n] i32, angle [n] f32,
    response [n] f32, size [n] f32, in OpenCV's order (level by level, row-major inside a level).  `exact_order` is False
    when a level held more corners than its quota: OpenCV then keeps every corner whose response reaches the quota-th
    largest one (reproduced) in an order that std::nth_element leaves behind (not reproduced: row-major here)."""
    h, w = img.shape
    raw = [np.ascontiguousarray(img, np.uint8)]
    for level in range(1, nlevels):
        dw, dh = level_size(w, h, level)
        raw.append(resize_linear_exact(raw[-1], dw, dh))
    quotas = level_quotas(nlevels)
    pts, octs, angs, resp, sizes, exact = [], [], [], [], [], True
    for level, im in enumerate(raw):
        lh, lw = im.shape
        score = fast_scores(im)
        ys, xs = np.nonzero(nonmax(score))
        inside = (xs >= EDGE_THRESHOLD) & (xs < lw - EDGE_THRESHOLD) & (ys >= EDGE_THRESHOLD) & (ys < lh - EDGE_THRESHOLD)
        ys, xs = ys[inside], xs[inside]
        r = score[ys, xs]
        if len(xs) > quotas[level]:  # KeyPointsFilter::retainBest
            exact = False
            thr = np.sort(r)[::-1][quotas[level] - 1] if quotas[level] > 0 else np.inf
            sel = r >= thr
            ys, xs, r = ys[sel], xs[sel], r[sel]
        sf = level_scale(level)
        pts.append(np.stack([xs.astype(np.float32) * sf, ys.astype(np.float32) * sf], 1).astype(np.float32))
        octs.append(np.full(len(xs), level, np.int32))
        angs.append(ic_angles(im, xs, ys) if len(xs) else np.zeros(0, np.float32))
        resp.append(r.astype(np.float32))
        sizes.append(np.full(len(xs), F32(PATCH_SIZE) * sf, np.float32))
    return {"pt": np.concatenate(pts), "octave": np.concatenate(octs), "angle": np.concatenate(angs),
            "response": np.concatenate(resp), "size": np.concatenate(sizes), "exact_order": exact}


def compute(img, pts, octaves, angles, pattern=None):
    """cv2.ORB.compute(img, keypoints) -> descriptors [n, 32] uint8 for keypoints given as (pt, octave, angle) arrays."""
    pattern = np.load(PATTERN_PATH) if pattern is None else pattern
    levels = pyramid(img)
    cx, cy, a, b = keypoint_params(pts, octaves, angles)
    p = pattern.astype(np.float32)
    out = np.zeros((len(cx), 32), np.uint8)
    for n in range(len(cx)):
        im = levels[int(octaves[n])]
        xa = p[:, 0] * a[n] - p[:, 1] * b[n]
        ya = p[:, 0] * b[n] + p[:, 1] * a[n]
        xb = p[:, 2] * a[n] - p[:, 3] * b[n]
        yb = p[:, 2] * b[n] + p[:, 3] * a[n]
        t0 = im[np.rint(ya).astype(np.int64) + cy[n], np.rint(xa).astype(np.int64) + cx[n]]
        t1 = im[np.rint(yb).astype(np.int64) + cy[n], np.rint(xb).astype(np.int64) + cx[n]]
        out[n] = np.packbits((t0 < t1).astype(np.uint8), bitorder="little")
    return out
