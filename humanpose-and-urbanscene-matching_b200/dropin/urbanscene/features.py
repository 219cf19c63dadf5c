"""urbanscene/features.py of the reference with the same signatures; kNN, ratio filter, RANSAC homography, inlier
gather and the corner projection run on the GPU (features.py:42,:45-55,:57-105,:177-234)."""
import logging
import os

import numpy as np
import torch

from .. import thresholds
from .._dev import engine, up, down

NORM_L2, NORM_HAMMING = 4, 6  # cv2's constants, so callers can keep passing them


class DMatch(object):
    """The fields of cv2.DMatch the reference reads (features.py:48-51)."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx, trainIdx, distance):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = int(queryIdx), int(trainIdx), 0, float(distance)


def _pad_binary(desc):
    """Binary descriptors zero-padded to the kernel's row widths (32 or 64 bytes): AKAZE's default MLDB descriptor is
    61 bytes wide.  Zero bytes on both sides leave every Hamming distance unchanged."""
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    if desc.ndim != 2:
        raise ValueError("knnMatch: descriptors must be a 2-D uint8 array")
    w = desc.shape[1]
    if w in (32, 64):
        return desc
    if w > 64:
        raise ValueError("knnMatch: binary descriptors wider than 64 bytes are not supported (got %d)" % w)
    out = np.zeros((desc.shape[0], 32 if w < 32 else 64), np.uint8)
    out[:, :w] = desc
    return out


class BFMatcher(object):
    """Duck-type of cv2.BFMatcher(norm) (crossCheck off) as features.match_and_draw uses it."""

    def __init__(self, normType=NORM_L2, seed=0):
        self.normType = normType
        self.last = None  # device tensors of the last knnMatch: (idx, dist) -- lets filter_matches stay on the GPU

    def knn2(self, queryDescriptors, trainDescriptors):
        if self.normType == NORM_HAMMING:
            q, t = _pad_binary(queryDescriptors), _pad_binary(trainDescriptors)
            idx, dist = engine.knn2_hamming(up(q, np.uint8), up(t, np.uint8))
        else:
            idx, dist = engine.knn2_l2(up(queryDescriptors, np.float32), up(trainDescriptors, np.float32))
        self.last = (idx, dist)
        return idx, dist

    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=2):
        if queryDescriptors is None or trainDescriptors is None:
            raise ValueError("knnMatch: descriptors are None (image without keypoints)")  # cv2 raises cv2.error here
        if k != 2:
            raise ValueError("this matcher implements the reference's call, knnMatch(..., k=2)")
        idx, dist = self.knn2(queryDescriptors, trainDescriptors)
        idx, dist = down(idx), down(dist).astype(np.float32)
        out = []
        for i in range(idx.shape[0]):
            out.append([DMatch(i, idx[i, j], dist[i, j]) for j in range(2) if idx[i, j] >= 0])
        return out


class KeyPointArray(object):
    """The keypoints of one image as arrays (what the device path produces), quacking like the list of cv2.KeyPoint the
    reference passes around: len(), indexing and iteration give objects with .pt, .size, .angle, .response, .octave,
    .class_id; `.xy` is the [n,2] float32 array the GPU matcher's gather takes directly."""

    class _KP(object):
        __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

        def __init__(self, pt, size, angle, response, octave):
            self.pt, self.size, self.angle, self.response, self.octave, self.class_id = pt, size, angle, response, octave, -1

    def __init__(self, xy, size, angle, response, octave):
        self.xy, self.size, self.angle, self.response, self.octave = xy, size, angle, response, octave

    def __len__(self):
        return len(self.xy)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return KeyPointArray(self.xy[i], self.size[i], self.angle[i], self.response[i], self.octave[i])
        return KeyPointArray._KP((float(self.xy[i, 0]), float(self.xy[i, 1])), float(self.size[i]), float(self.angle[i]),
                                 float(self.response[i]), int(self.octave[i]))

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def to_cv2(self):
        import cv2
        return [cv2.KeyPoint(float(x), float(y), float(s), float(a), float(r), int(o), -1)
                for (x, y), s, a, r, o in zip(self.xy, self.size, self.angle, self.response, self.octave)]


class OrbGpu(object):
    """The reference's ORB (features.py:19-20) on the GPU: `detectAndCompute(img, None)` returns cv2's keypoints (same
    order, identical pt / octave / size / response / angle, as a KeyPointArray) and cv2's descriptors bit for bit --
    pyramid, FAST + non-maximum suppression, orientation, smoothing and the rBRIEF tests all run on the device
    (hpusm_orb_detect_and_compute).  There is NO silent CPU path: a detection mask (the reference never passes one) and an
    image dense enough for OpenCV's per-level keypoint cap to bite (its cull is a partial sort whose tie order is not
    specified, so it cannot be reproduced bit for bit) raise, with the two explicit ways out in the message --
    HPUSM_ORB_CPU_DETECT=1 (OpenCV detects on the CPU for exactly those calls, description stays on the GPU) or
    HPUSM_ORB=cv2 (OpenCV's ORB altogether).  `last_descriptors` keeps the device tensor of the last call so a caller that
    goes on to match can skip the upload."""

    def __init__(self, cv_detector):
        self.cv = cv_detector
        self.last_descriptors = None

    def _cpu_detect(self, image, mask, why):
        if os.environ.get("HPUSM_ORB_CPU_DETECT", "") != "1":
            raise NotImplementedError("OrbGpu: %s is not done on the device; set HPUSM_ORB_CPU_DETECT=1 to let OpenCV detect on "
                                      "the CPU for such calls (description stays on the GPU), or HPUSM_ORB=cv2 for OpenCV's ORB" % why)
        logging.warning("OrbGpu: %s: detection by OpenCV on the CPU (HPUSM_ORB_CPU_DETECT=1)", why)
        return self.cv.detect(image, mask)

    @staticmethod
    def _gray(image):
        img = np.ascontiguousarray(image)
        if img.ndim != 2 or img.dtype != np.uint8:
            raise ValueError("OrbGpu: grayscale uint8 images only (the reference reads them with cv2.imread(path, 0))")
        return img

    def detect(self, image, mask=None):
        if mask is not None:
            return self._cpu_detect(image, mask, "detection under a mask")
        return self.detectAndCompute(image, None, want_descriptors=False)[0]

    def compute(self, image, keypoints):
        if len(keypoints) == 0:
            self.last_descriptors = None
            return keypoints, None
        if isinstance(keypoints, KeyPointArray):
            pts, octs, angs = keypoints.xy, keypoints.octave, keypoints.angle
        else:
            pts = np.float32([kp.pt for kp in keypoints])
            octs = np.int32([kp.octave for kp in keypoints])
            angs = np.float32([kp.angle for kp in keypoints])
        lxy, cs = engine.orb_keypoint_params(pts, octs, angs)
        desc = engine.orb_compute(up(self._gray(image), np.uint8), up(lxy, np.int32), up(cs, np.float32))
        self.last_descriptors = desc
        return keypoints, down(desc)

    def detectAndCompute(self, image, mask=None, want_descriptors=True):
        if mask is not None:
            return self.compute(image, self._cpu_detect(image, mask, "detection under a mask"))
        out = engine.orb_detect_and_compute(up(self._gray(image), np.uint8), want_descriptors=want_descriptors)
        if out["over_quota"]:  # OpenCV would cull this level by response (retainBest): not reproducible bit for bit
            kps = self._cpu_detect(image, None, "a pyramid level with more keypoints than OpenCV's per-level quota")
            return self.compute(image, kps) if want_descriptors else (kps, None)
        kps = KeyPointArray(down(out["pt"]), down(out["size"]), down(out["angle"]), down(out["response"]), down(out["octave"]))
        self.last_descriptors = out["desc"]
        return kps, (down(out["desc"]) if want_descriptors else None)

    def __getattr__(self, name):  # everything else (getters, setters) is the OpenCV detector's
        return getattr(self.cv, name)


def init_feature(name):
    """(detector, matcher) as features.py:10-43.  Detection stays with OpenCV (out of the path's scope); the matcher
    is ours.  The '-flann' names (every driver script of the reference passes 'orb-flann': urbanscene/main.py:15,
    jochen_deel.py:14, dataset/dataset_urbanscene.py:11) get the same exact brute-force GPU matcher: FLANN's KD-tree / LSH
    index only approximates the k=2 neighbours that this matcher returns exactly, and match_and_draw consumes both through
    the same knnMatch contract.  Set HPUSM_FLANN=cv2 to get the reference's own cv2.FlannBasedMatcher instead."""
    import cv2
    chunks = name.split('-')
    if chunks[0] == 'sift':
        detector = cv2.SIFT_create() if hasattr(cv2, "SIFT_create") else cv2.xfeatures2d.SIFT_create()
        norm = NORM_L2
    elif chunks[0] == 'surf':
        detector = cv2.xfeatures2d.SURF_create(800)
        norm = NORM_L2
    elif chunks[0] == 'orb':
        detector = cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                                  scoreType=cv2.ORB_FAST_SCORE, patchSize=31)
        if os.environ.get("HPUSM_ORB", "gpu").lower() != "cv2":
            detector = OrbGpu(detector)  # same keypoints, same descriptors bit for bit; description on the GPU
        norm = NORM_HAMMING
    elif chunks[0] == 'akaze':
        detector = cv2.AKAZE_create()
        norm = NORM_HAMMING
    elif chunks[0] == 'brisk':
        detector = cv2.BRISK_create()
        norm = NORM_HAMMING
    else:
        return None, None
    if 'flann' in chunks:
        if os.environ.get("HPUSM_FLANN", "").lower() == "cv2":  # the reference's approximate matcher, features.py:30-40
            search_params = dict(checks=50)
            if norm == NORM_L2:
                flann_params = dict(algorithm=thresholds.FLANN_INDEX_KDTREE, trees=5)
            else:
                flann_params = dict(algorithm=thresholds.FLANN_INDEX_LSH, table_number=6, key_size=12, multi_probe_level=1)
            return detector, cv2.FlannBasedMatcher(flann_params, search_params)
        logging.info("init_feature(%r): FLANN requested, using the exact GPU brute-force matcher", name)
    return detector, BFMatcher(norm)


def _pt(kp):
    return kp.pt if hasattr(kp, "pt") else (float(kp[0]), float(kp[1]))


def filter_matches(kp1, kp2, matches, ratio=thresholds.FILTER_RATIO):
    mkp1, mkp2 = [], []
    for m in matches:
        if len(m) == 2 and m[0].distance < m[1].distance * ratio:
            m = m[0]
            mkp1.append(kp1[m.queryIdx])
            mkp2.append(kp2[m.trainIdx])
    p1 = np.float32([_pt(kp) for kp in mkp1]).reshape(-1, 1, 2)
    p2 = np.float32([_pt(kp) for kp in mkp2]).reshape(-1, 1, 2)
    return p1, p2, list(zip(mkp1, mkp2))


def find_homography_ransac(src, dst, thresh=5.0, maxIters=2000, confidence=0.995, seed=0, sampler="cv2"):
    """cv2.findHomography(src, dst, cv2.RANSAC, thresh) -> (H float64 (3,3) or None, mask uint8 (K,1)).

    The default sampler replays OpenCV's own (fixed-seed) sample sequence on the GPU, so H and mask are the ones
    cv2.findHomography returns; sampler="counter" draws from the seeded counter-based table instead."""
    src = np.float32(src).reshape(-1, 2)
    dst = np.float32(dst).reshape(-1, 2)
    dsrc = up(src, np.float32)
    offs = torch.tensor([0, len(src)], dtype=torch.int64, device=dsrc.device)
    H, mask, ninl, best = engine.ransac_homography_batch(dsrc, up(dst, np.float32), offs, thresh, maxIters, confidence, seed,
                                                         sampler=sampler)
    if int(down(best)[0]) < 0:
        return None, None
    return down(H)[0], down(mask).reshape(-1, 1)


def find_homography_both_ways(p_a, p_b, thresh=5.0, maxIters=2000, confidence=0.995, seed=0, sampler="cv2"):
    """The two calls of match_and_draw (features.py:66 and :69) as ONE batch of two pairs: (H, mask, H2, mask2)."""
    a = np.float32(p_a).reshape(-1, 2)
    b = np.float32(p_b).reshape(-1, 2)
    k = len(a)
    src = up(np.concatenate([a, b]), np.float32)
    dst = up(np.concatenate([b, a]), np.float32)
    offs = torch.tensor([0, k, 2 * k], dtype=torch.int64, device=src.device)
    H, mask, ninl, best = engine.ransac_homography_batch(src, dst, offs, thresh, maxIters, confidence, seed, sampler=sampler)
    best, H, mask = down(best), down(H), down(mask)
    out = []
    for p in range(2):
        out += [None, None] if best[p] < 0 else [H[p], mask[p * k:(p + 1) * k].reshape(-1, 1)]
    return tuple(out)


def _kp_array(kps):
    if isinstance(kps, KeyPointArray):
        return np.ascontiguousarray(kps.xy, dtype=np.float32).reshape(-1, 2)
    if isinstance(kps, np.ndarray):
        return np.ascontiguousarray(kps, dtype=np.float32).reshape(-1, 2)
    return np.float32([_pt(kp) for kp in kps]).reshape(-1, 2)


def _match_and_filter_on_device(matcher, desc_model, desc_input, kp_model, kp_input, ratio):
    """knnMatch + filter_matches without materialising 2 x Nq DMatch objects: kNN-2, ratio test, compaction and
    keypoint gather stay on the GPU; only the K surviving point pairs come back."""
    idx, dist = matcher.knn2(desc_model, desc_input)
    keep, count = engine.ratio_filter(idx, dist, ratio)
    p_q, p_t, sel, cnt = engine.gather_matches(up(_kp_array(kp_model), np.float32), up(_kp_array(kp_input), np.float32), idx, keep)
    n = int(cnt.item())
    return down(p_q[:n]).reshape(-1, 1, 2), down(p_t[:n]).reshape(-1, 1, 2)


def match_and_draw(win, matcher, desc_model, desc_input, kp_model, kp_input, model_img, input_img, show_win):
    if isinstance(matcher, BFMatcher):
        if desc_model is None or desc_input is None:
            raise ValueError("knnMatch: descriptors are None (image without keypoints)")
        p_model, p_input = _match_and_filter_on_device(matcher, desc_model, desc_input, kp_model, kp_input,
                                                       thresholds.FILTER_RATIO)
    else:  # any other object with cv2's knnMatch contract
        raw_matches = matcher.knnMatch(desc_model, trainDescriptors=desc_input, k=2)
        p_model, p_input, kp_pairs = filter_matches(kp_model, kp_input, raw_matches)
    if len(p_model) >= thresholds.MIN_MATCH_COUNT:
        H, mask, H2, mask2 = find_homography_both_ways(p_model, p_input, 5.0)
        if H is None or H2 is None:
            return (None, None, None, None, None, None)
        logging.debug('%d / %d  inliers/matched' % (np.sum(mask), len(mask)))
        h, w = model_img.shape
        homography_pts = np.float32([[0, 0], [0, h - 1], [w - 1, h - 1], [w - 1, 0]]).reshape(-1, 1, 2)
        homography_dst = down(engine.perspective_transform(up(homography_pts, np.float32), up(H, np.float64)))
        try:  # the reference draws the projected model outline INTO input_img (callers rely on that side effect)
            import cv2
            cv2.polylines(input_img, [np.int32(homography_dst)], True, 255, 3, cv2.LINE_AA)
        except ImportError:
            pass
        my_mask = np.asarray(mask).astype(bool)
        good_model_pts = np.squeeze(p_model[np.array(my_mask)][:])
        good_input_pts = np.squeeze(p_input[np.array(my_mask)][:])
    else:
        H, mask = None, None
        logging.debug('%d matches found, not enough for homography estimation' % len(p_model))
        return (mask, None, None, H, None, None)
    return (mask, good_model_pts, good_input_pts, H, H2, len(good_model_pts))


def validate_homography(perspective_trans_matrix):
    H = perspective_trans_matrix
    if H[0, 0] * H[1, 1] - H[0, 1] * H[1, 0] < 0:
        return False
    N1 = (H[0, 0] * H[0, 0] + H[0, 1] * H[0, 1]) ** 0.5
    if N1 > 4 or N1 < 0.1:
        return False
    N2 = (H[1, 0] * H[1, 0] + H[1, 1] * H[1, 1]) ** 0.5
    if N2 > 4 or N2 < 0.1:
        return False
    N3 = (H[2, 0] * H[2, 0] + H[2, 1] * H[2, 1]) ** 0.5
    if N3 > 0.002:
        return False
    return True


def euclidean_distance(model, transformed_input):
    d = np.abs(model - transformed_input)
    return ((d[:, 0]) ** 2 + d[:, 1] ** 2) ** 0.5


def max_euclidean_distance(model, transformed_input):
    return max(euclidean_distance(model, transformed_input))
