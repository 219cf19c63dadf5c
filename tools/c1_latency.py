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
print(json.dumps(out))
