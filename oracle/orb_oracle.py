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
NLEVELS, SCALE_FACTOR = 8, 1.2


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
