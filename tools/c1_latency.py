#!/usr/bin/env python
"""Config C1 (pisa1 x pisa2, the reference's own CPU-runnable case): wall time of features.match_and_draw through the
drop-in (GPU) next to the same calls on cv2 (CPU), on the descriptors/keypoints recorded in tests/golden/."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import torch  # noqa: E402

ft = importlib.import_module("humanpose-and-urbanscene-matching_b200.dropin.urbanscene.features")


def best_of(fn, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


out = {}
for det in ("sift", "orb"):
    g = np.load(os.path.join(ROOT, "tests", "golden", "scene_pisa1_pisa2_%s.npz" % det))
    dm = g["desc_model"].astype(np.float32) if det == "sift" else g["desc_model"]
    di = g["desc_input"].astype(np.float32) if det == "sift" else g["desc_input"]
    kp_m, kp_i = g["kp_model"], g["kp_input"]
    model_img, input_img = np.zeros(tuple(g["model_shape"]), np.uint8), np.zeros(tuple(g["input_shape"]), np.uint8)
    ours = ft.BFMatcher(ft.NORM_L2 if det == "sift" else ft.NORM_HAMMING)
    ref = cv2.BFMatcher(cv2.NORM_L2 if det == "sift" else cv2.NORM_HAMMING)

    def run_ours():
        return ft.match_and_draw("w", ours, dm, di, kp_m, kp_i, model_img, input_img, False)

    def run_ours_knn():
        return ours.knn2(dm, di)

    def run_cv2():
        raw = ref.knnMatch(dm, trainDescriptors=di, k=2)
        p1, p2, _ = ft.filter_matches(kp_m, kp_i, raw)
        H, mask = cv2.findHomography(p1, p2, cv2.RANSAC, 5.0)
        H2, mask2 = cv2.findHomography(p2, p1, cv2.RANSAC, 5.0)
        return H, H2

    def run_cv2_knn():
        return ref.knnMatch(dm, trainDescriptors=di, k=2)

    run_ours(); run_cv2()
    out[det] = {"shape": "%dx%d" % (len(dm), len(di)), "ours_match_and_draw_ms": best_of(run_ours), "ours_knn_only_ms": best_of(run_ours_knn),
                "cv2_match_and_draw_ms": best_of(run_cv2), "cv2_knn_only_ms": best_of(run_cv2_knn), "cv2_threads": cv2.getNumThreads()}
# feature extraction (urban_scene.py:31-32: detector.detectAndCompute), the step in front of the matcher
g = np.load(os.path.join(ROOT, "tests", "golden", "orb_descriptors.npz"))
ours_det, _ = ft.init_feature("orb")
ref_det = cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_FAST_SCORE, patchSize=31)
for k, name in enumerate(g["names"]):
    img = g["img_%d" % k]
    dimg = torch.from_numpy(img).cuda()
    hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
    ours_det.detectAndCompute(img, None); ref_det.detectAndCompute(img, None)
    out["orb_extract_%s" % str(name)] = {
        "shape": "%dx%d" % img.shape, "keypoints": int(len(g["pt_%d" % k])),
        "ours_detectAndCompute_ms": best_of(lambda: ours_det.detectAndCompute(img, None)),
        "ours_device_only_ms": best_of(lambda: hp.engine.orb_detect_and_compute(dimg)),
        "cv2_detectAndCompute_ms": best_of(lambda: ref_det.detectAndCompute(img, None))}
print(json.dumps(out))
