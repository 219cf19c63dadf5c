"""Records the golden vectors under tests/golden/ from the UNMODIFIED reference -- test infrastructure only.

Run in the build container (needs /root/reference; the GPU box never runs this):
    python oracle/record_golden.py

The reference ships no tests and no expected outputs, so these files ARE the pins:
every array below is the output of the reference's own modules (imported from /root/reference with
the three-line shim of SURVEY 8c: inert matplotlib, numpy.math, cv2.SIFT_create) or of the OpenCV /
NumPy binaries it calls at the cited call sites, on the shipped fixtures (img/*.jpg, json_data/*.json)
and on seeded synthetic inputs.  Versions are stored in every file.
"""
import glob
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HPUSM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)
np.math = math  # pose_comparison.py:12 uses np.math, removed in numpy 2

import cv2  # noqa: E402
import common as ref_common  # noqa: E402
import posematching.pose_comparison as ref_pc  # noqa: E402
import posematching.procrustes as ref_procrustes  # noqa: E402
import posematching.single_person as ref_sp  # noqa: E402
import urbanscene.features as ref_features  # noqa: E402

sys.path.insert(0, ROOT)
import importlib  # noqa: E402
synth = importlib.import_module("humanpose-and-urbanscene-matching_b200.synth")

VERSIONS = np.array(["opencv %s" % cv2.__version__, "numpy %s" % np.__version__])


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, versions=VERSIONS, **arrays)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def matches_to_arrays(raw, nq):
    idx = -np.ones((nq, 2), np.int32)
    dist = np.zeros((nq, 2), np.float32)
    for i, m in enumerate(raw):
        for j, mm in enumerate(m[:2]):
            idx[i, j] = mm.trainIdx
            dist[i, j] = mm.distance
            assert mm.queryIdx == i
    return idx, dist


def record_scene_pair(tag, img_model, img_input):
    """C1: features.match_and_draw on a real image pair, SIFT (L2) and ORB (Hamming), BF matcher."""
    model = cv2.imread(os.path.join(REF, "img", img_model), 0)
    inp = cv2.imread(os.path.join(REF, "img", img_input), 0)
    for det_name in ("sift", "orb"):
        if det_name == "sift":
            detector, matcher = cv2.SIFT_create(), cv2.BFMatcher(cv2.NORM_L2)  # features.py:13 no longer exists
        else:
            detector, matcher = ref_features.init_feature("orb")
        kp_m, desc_m = detector.detectAndCompute(model, None)
        kp_i, desc_i = detector.detectAndCompute(inp, None)
        raw = matcher.knnMatch(desc_m, trainDescriptors=desc_i, k=2)  # features.py:59
        idx, dist = matches_to_arrays(raw, len(kp_m))
        p1, p2, _ = ref_features.filter_matches(kp_m, kp_i, raw)  # features.py:45
        out = ref_features.match_and_draw("w", matcher, desc_m, desc_i, kp_m, kp_i, model, inp.copy(), False)
        mask, good_m, good_i, H, H2, n_good = out
        arrays = dict(
            desc_model=desc_m.astype(np.uint8) if det_name == "sift" else desc_m,
            desc_input=desc_i.astype(np.uint8) if det_name == "sift" else desc_i,
            kp_model=np.float32([k.pt for k in kp_m]), kp_input=np.float32([k.pt for k in kp_i]),
            model_shape=np.array(model.shape), input_shape=np.array(inp.shape),
            knn_idx=idx, knn_dist=dist, p_model=p1, p_input=p2)
        if det_name == "sift":
            assert np.array_equal(desc_m, np.rint(desc_m)) and desc_m.max() <= 255
        if H is not None:
            arrays.update(mask=mask, good_model_pts=good_m, good_input_pts=good_i, H=H, H2=H2, n_good=np.array(n_good),
                          valid=np.array(ref_features.validate_homography(H)))
        save("scene_%s_%s.npz" % (tag, det_name), **arrays)


EXTRA_SIFT_PAIRS = [("taj1_taj2", "taj1.jpg", "taj2.jpg"), ("eifel2_eifel3", "eifel2.jpg", "eifel3.jpg"),
                    ("india1_india2", "india1.jpg", "india2.jpg")]


def record_scene_pairs_sift_extra():
    """More real image pairs of the reference's img/ directory, SIFT + BFMatcher(NORM_L2).knnMatch (features.py:59) +
    filter_matches (features.py:45): real descriptor statistics (not the synthetic generator's) for the kNN engines."""
    for tag, img_model, img_input in EXTRA_SIFT_PAIRS:
        model = cv2.imread(os.path.join(REF, "img", img_model), 0)
        inp = cv2.imread(os.path.join(REF, "img", img_input), 0)
        detector, matcher = cv2.SIFT_create(), cv2.BFMatcher(cv2.NORM_L2)
        kp_m, desc_m = detector.detectAndCompute(model, None)
        kp_i, desc_i = detector.detectAndCompute(inp, None)
        raw = matcher.knnMatch(desc_m, trainDescriptors=desc_i, k=2)
        idx, dist = matches_to_arrays(raw, len(kp_m))
        p1, p2, _ = ref_features.filter_matches(kp_m, kp_i, raw)
        assert np.array_equal(desc_m, np.rint(desc_m)) and desc_m.max() <= 255 and desc_m.min() >= 0
        save("scene_%s_sift_knn.npz" % tag, desc_model=desc_m.astype(np.uint8), desc_input=desc_i.astype(np.uint8),
             kp_model=np.float32([k.pt for k in kp_m]), kp_input=np.float32([k.pt for k in kp_i]), knn_idx=idx, knn_dist=dist,
             p_model=p1, p_input=p2)


def record_scene_pair_orb_extra():
    """One more real image pair through the reference's ORB detector + BFMatcher(NORM_HAMMING).knnMatch + filter_matches."""
    for tag, img_model, img_input in [("taj1_taj2", "taj1.jpg", "taj2.jpg")]:
        model = cv2.imread(os.path.join(REF, "img", img_model), 0)
        inp = cv2.imread(os.path.join(REF, "img", img_input), 0)
        detector, matcher = ref_features.init_feature("orb")
        kp_m, desc_m = detector.detectAndCompute(model, None)
        kp_i, desc_i = detector.detectAndCompute(inp, None)
        raw = matcher.knnMatch(desc_m, trainDescriptors=desc_i, k=2)
        idx, dist = matches_to_arrays(raw, len(kp_m))
        p1, p2, _ = ref_features.filter_matches(kp_m, kp_i, raw)
        save("scene_%s_orb_knn.npz" % tag, desc_model=desc_m, desc_input=desc_i, kp_model=np.float32([k.pt for k in kp_m]),
             kp_input=np.float32([k.pt for k in kp_i]), knn_idx=idx, knn_dist=dist, p_model=p1, p_input=p2)


def record_knn_synthetic():
    # L2, SIFT-like integer-valued descriptors (C2 generator, small), with planted matches and exact ties
    q, t = synth.sift_like(700, 1500, seed=10, planted=0.1)
    t[7] = t[3]  # duplicate rows -> distance ties that must resolve to the lower index
    t[1400] = t[9]
    raw = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, trainDescriptors=t, k=2)
    idx, dist = matches_to_arrays(raw, len(q))
    save("knn_l2_siftlike.npz", q=q.astype(np.uint8), t=t.astype(np.uint8), idx=idx, dist=dist)
    # L2, general float descriptors
    rng = np.random.default_rng(11)
    q = rng.standard_normal((300, 128)).astype(np.float32)
    t = rng.standard_normal((900, 128)).astype(np.float32)
    t[100:130] = q[:30] + 0.05 * rng.standard_normal((30, 128)).astype(np.float32)
    raw = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, trainDescriptors=t, k=2)
    idx, dist = matches_to_arrays(raw, len(q))
    save("knn_l2_float.npz", q=q, t=t, idx=idx, dist=dist)
    # L2, odd width (64-d, SURF-like) and a single train row
    q = rng.random((50, 64)).astype(np.float32)
    t = rng.random((1, 64)).astype(np.float32)
    raw = cv2.BFMatcher(cv2.NORM_L2).knnMatch(q, trainDescriptors=t, k=2)
    idx, dist = matches_to_arrays(raw, len(q))
    save("knn_l2_single_train.npz", q=q, t=t, idx=idx, dist=dist)
    # Hamming, 256-bit, with planted near-duplicates (many ties by construction)
    q, t = synth.orb_like(600, 2100, seed=12, planted=0.2)
    raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, trainDescriptors=t, k=2)
    idx, dist = matches_to_arrays(raw, len(q))
    save("knn_hamming_256.npz", q=q, t=t, idx=idx.astype(np.int32), dist=dist.astype(np.int32))
    # Hamming, 512-bit (BRISK width)
    rng = np.random.default_rng(13)
    q = rng.integers(0, 256, (200, 64), dtype=np.uint8)
    t = rng.integers(0, 256, (777, 64), dtype=np.uint8)
    raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, trainDescriptors=t, k=2)
    idx, dist = matches_to_arrays(raw, len(q))
    save("knn_hamming_512.npz", q=q, t=t, idx=idx.astype(np.int32), dist=dist.astype(np.int32))


def record_ransac_cv2():
    """cv2.findHomography on inputs with a unique consensus set (features.py:66): pins H and mask."""
    srcs, dsts, Hs, masks, offs = [], [], [], [], [0]
    cases = [(400, 0.6, 1.0), (2000, 0.3, 1.0), (161, 0.5, 1.0), (64, 0.8, 1.0), (1000, 0.45, 1.0),
             (400, 0.6, 0.0), (2000, 0.3, 0.0), (161, 0.5, 0.0), (16, 0.9, 0.0), (1000, 0.45, 0.0)]
    for k, (n, ratio, noise) in enumerate(cases):
        src, dst, _ = synth.ransac_pair(n, inlier_ratio=ratio, seed=20 + k, noise=noise)
        H, mask = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 5.0)
        srcs.append(src); dsts.append(dst); Hs.append(H); masks.append(mask.ravel()); offs.append(offs[-1] + n)
    save("ransac_cv2.npz", src=np.concatenate(srcs), dst=np.concatenate(dsts), H=np.stack(Hs),
         mask=np.concatenate(masks), pair_offsets=np.array(offs, np.int64), noise=np.array([c[2] for c in cases]))
    # cv2.perspectiveTransform (features.py:77, transformation.py:38)
    pts = np.random.default_rng(29).uniform(0, 640, (257, 1, 2)).astype(np.float32)
    save("perspective_transform.npz", pts=pts, H=Hs[0], out=cv2.perspectiveTransform(pts, Hs[0]))


RANSAC_REAL_PAIRS = [("sift", "pisa9.jpg", "pisa10.jpg"), ("sift", "notredam1.jpg", "notredam2.jpg"),
                     ("sift", "india1.jpg", "india2.jpg"), ("sift", "jochen_foto4.jpg", "jochen_foto5.jpg"),
                     ("sift", "foto3.jpg", "foto4.jpg"), ("sift", "emma1.jpg", "emma2.jpg"),
                     ("orb", "taj1.jpg", "taj3.jpg"), ("orb", "eifel2.jpg", "eifel4.jpg"),
                     ("orb", "notredam1.jpg", "notredam3.jpg"), ("orb", "pisa9.jpg", "pisa10.jpg"),
                     # pairs whose homography the reference itself rejects (validate_homography False): ill-posed fits
                     ("sift", "pisa1.jpg", "pisa2.jpg"), ("orb", "pisa1.jpg", "pisa2.jpg"), ("sift", "pisa8.jpg", "pisa9.jpg")]


def record_ransac_real_pairs():
    """features.match_and_draw (features.py:57-105) on real image pairs of the reference's img/ directory: the matched
    point lists (filter_matches, features.py:45) and what the two cv2.findHomography calls (features.py:66,:69) return.
    Ten pairs have a homography the reference accepts (validate_homography, features.py:177), three do not."""
    arrays, names = {}, []
    for k, (det_name, img_model, img_input) in enumerate(RANSAC_REAL_PAIRS):
        model = cv2.imread(os.path.join(REF, "img", img_model), 0)
        inp = cv2.imread(os.path.join(REF, "img", img_input), 0)
        if det_name == "sift":
            detector, matcher = cv2.SIFT_create(), cv2.BFMatcher(cv2.NORM_L2)
        else:
            detector, matcher = ref_features.init_feature("orb")
        kp_m, desc_m = detector.detectAndCompute(model, None)
        kp_i, desc_i = detector.detectAndCompute(inp, None)
        raw = matcher.knnMatch(desc_m, trainDescriptors=desc_i, k=2)
        p1, p2, _ = ref_features.filter_matches(kp_m, kp_i, raw)
        mask, good_m, good_i, H, H2, n_good = ref_features.match_and_draw("w", matcher, desc_m, desc_i, kp_m, kp_i, model,
                                                                           inp.copy(), False)
        H2b, mask2 = cv2.findHomography(p2, p1, cv2.RANSAC, 5.0)  # features.py:69 (the reference drops this mask)
        assert np.array_equal(H2, H2b)
        names.append("%s:%s:%s" % (det_name, img_model, img_input))
        arrays.update({"p_model_%d" % k: p1, "p_input_%d" % k: p2, "mask_%d" % k: mask, "mask2_%d" % k: mask2,
                       "H_%d" % k: H, "H2_%d" % k: H2, "n_good_%d" % k: np.array(n_good),
                       "valid_%d" % k: np.array([ref_features.validate_homography(H), ref_features.validate_homography(H2)])})
    save("ransac_real_pairs.npz", names=np.array(names), **arrays)


def record_ransac_synth_cv2():
    """cv2.findHomography(src, dst, cv2.RANSAC, 5.0) (features.py:66) on 1000 seeded noisy synthetic pairs: H, mask."""
    Hs, masks, offs = [], [], [0]
    from oracle.ransac_cases import N_SYNTH, ransac_synth_case
    for k in range(N_SYNTH):
        src, dst, _ = ransac_synth_case(k)
        H, mask = cv2.findHomography(src.reshape(-1, 1, 2), dst.reshape(-1, 1, 2), cv2.RANSAC, 5.0)
        assert H is not None
        Hs.append(H); masks.append(mask.ravel()); offs.append(offs[-1] + len(src))
    save("ransac_synth_cv2.npz", H=np.stack(Hs), mask_bits=np.packbits(np.concatenate(masks)),
         pair_offsets=np.array(offs, np.int64))


def record_validate_homography():
    """features.validate_homography (features.py:177-211) on 1000 matrices: random similarity/projective ones around every
    threshold (determinant sign, row norms 0.1 / 4, perspective norm 0.002), exactly singular blocks, and edge values."""
    rng = np.random.default_rng(77)
    Hs = []
    for k in range(1000):
        kind = k % 8
        th = rng.uniform(-np.pi, np.pi)
        s1, s2 = rng.uniform(0.05, 5.0, 2) if kind != 1 else rng.choice([0.1, 4.0, 0.0999999, 4.0000001, 0.1000001, 3.9999999], 2)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        A = np.diag([s1, s2]) @ R
        if kind == 2:
            A[1] = -A[1]                                     # orientation reversing
        if kind == 3:
            A[1] = A[0] * rng.uniform(0.5, 2.0)              # singular top-left block (determinant ~ 0 of either sign)
        H = np.eye(3)
        H[:2, :2] = A
        H[:2, 2] = rng.uniform(-300, 300, 2)
        pn = rng.choice([0.0, 0.002, 0.0019999, 0.0020001, rng.uniform(0, 0.004)])
        ph = rng.uniform(-np.pi, np.pi)
        H[2, :2] = pn * np.array([np.cos(ph), np.sin(ph)])
        if kind == 4:
            H[:2, :2] = np.array([[1.0, 2.0], [2.0, 4.0]]) * rng.uniform(0.2, 0.9)  # exactly rank one
        if kind == 5:
            H[:2, :2] = rng.uniform(-4.5, 4.5, (2, 2))
        Hs.append(H)
    Hs = np.stack(Hs)
    valid = np.array([bool(ref_features.validate_homography(H)) for H in Hs])
    assert 50 < valid.sum() < 950
    save("validate_homography.npz", H=Hs, valid=valid)


def _reference_orb():
    detector, _ = ref_features.init_feature("orb")  # features.py:19-20: nfeatures=100000, 1.2, 8 levels, WTA_K=2, FAST score
    return detector


def record_orb_pattern():
    """OpenCV's 256-pair ORB test pattern, read off the binary: one keypoint at angle 0 on a black image with a single
    pixel of value 11 (the 7x7 Gaussian leaves exactly one pixel of value 1) sets the bits whose SECOND point is that
    pixel; white image / pixel 244 gives the FIRST points.  Saved as [256, 4] = (x0, y0, x1, y1)."""
    orb = _reference_orb()
    A = np.full((256, 2), 99, np.int32)
    B = np.full((256, 2), 99, np.int32)
    for py in range(-20, 21):
        for px in range(-20, 21):
            for inv in (False, True):
                im = np.full((200, 200), 255, np.uint8) if inv else np.zeros((200, 200), np.uint8)
                im[100 + py, 100 + px] = 244 if inv else 11
                _, d = orb.compute(im, [cv2.KeyPoint(100.0, 100.0, 31.0, 0.0, 1.0, 0, -1)])
                for i in np.nonzero(np.unpackbits(d[0], bitorder="little"))[0]:
                    tgt = A if inv else B
                    assert tgt[i, 0] == 99
                    tgt[i] = (px, py)
    assert (A != 99).all() and (B != 99).all()
    np.save(os.path.join(OUT, "orb_pattern.npy"), np.concatenate([A, B], 1).astype(np.int8))
    print("orb_pattern.npy", np.abs(A).max(), np.abs(B).max())


ORB_FIXTURES = ["pisa7.jpg", "taj1.jpg", "pisa2.jpg"]


def record_orb_descriptors():
    """detector.detectAndCompute (urban_scene.py:31-32) with the reference's ORB on three img/ fixtures: the grayscale image
    (the GPU box has no /root/reference), every keypoint's (pt, octave, angle) and its descriptor."""
    orb = _reference_orb()
    arrays = {}
    for k, name in enumerate(ORB_FIXTURES):
        img = cv2.imread(os.path.join(REF, "img", name), 0)
        kp, desc = orb.detectAndCompute(img, None)
        kp2, desc2 = orb.compute(img, kp)
        assert np.array_equal(desc, desc2) and len(kp2) == len(kp)
        kd = orb.detect(img, None)  # the same keypoints, through the detection-only entry point
        assert [(q.pt, q.octave, q.angle) for q in kd] == [(q.pt, q.octave, q.angle) for q in kp]
        arrays.update({"img_%d" % k: img, "pt_%d" % k: np.float32([q.pt for q in kp]),
                       "octave_%d" % k: np.int32([q.octave for q in kp]), "angle_%d" % k: np.float32([q.angle for q in kp]),
                       "response_%d" % k: np.float32([q.response for q in kp]), "size_%d" % k: np.float32([q.size for q in kp]),
                       "desc_%d" % k: desc})
    save("orb_descriptors.npz", names=np.array(ORB_FIXTURES), **arrays)


def load_fixture_poses():
    poses, names = [], []
    for f in sorted(glob.glob(os.path.join(REF, "json_data", "*.json"))):
        try:
            people = ref_common.parse_JSON_multi_person(f)  # common.py:215
        except Exception:
            continue
        for k, p in enumerate(people):
            if p.shape == (18, 2):
                poses.append(p)
                names.append("%s#%d" % (os.path.basename(f), k))
    return np.stack(poses), names


def record_pose():
    poses, names = load_fixture_poses()
    print("fixture poses:", poses.shape)
    rng = np.random.default_rng(30)
    # --- match_single on fixture pairs + synthetic perturbations of fixture poses
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, len(poses), (300, 2))]
    models = [poses[a] for a, _ in pairs]
    inputs = [poses[b] for _, b in pairs]
    db = synth.pose_database(poses, 500, seed=31)
    for k in range(len(db)):
        models.append(poses[k % 7])
        inputs.append(db[k].astype(np.float64))
    M, I = np.stack(models), np.stack(inputs)
    match = np.zeros(len(M), np.int8)
    score = np.full(len(M), np.nan)
    transf = np.full((len(M), 22, 2), np.nan)
    for k in range(len(M)):
        try:
            r = ref_sp.match_single(M[k], I[k], True)  # single_person.py:80
            match[k], score[k], transf[k] = int(r.match_bool), r.error_score, r.input_transformation
        except (ValueError, np.linalg.LinAlgError):
            match[k] = -1  # reference raises: np.min of an empty array (no detected joint) / NaN into LAPACK
    save("pose_match_single.npz", model=M, input=I, match=match, score=score, input_transformation=transf,
         fixture_poses=poses)
    # --- find_transformation on body parts, incl. rank-deficient parts (few detected joints)
    fm, fi, fo, fA, fh = [], [], [], [], []
    for k in range(200):
        a, b = poses[rng.integers(len(poses))], poses[rng.integers(len(poses))]
        rows = [ref_common.split_in_face_legs_torso_v2(x) for x in (a, b)]
        part = k % 3
        m, i = rows[0][part].copy(), rows[1][part].copy()
        if k % 5 == 0:  # leave only 0..2 detected joints
            keep = rng.integers(0, 3)
            i[keep:] = 0
        out, A = ref_common.find_transformation(m, i)  # common.py:37
        pad = np.zeros((10, 2))
        mm, ii, oo = pad.copy(), pad.copy(), pad.copy()
        mm[:len(m)], ii[:len(i)], oo[:len(i)] = m, i, out
        fm.append(mm); fi.append(ii); fo.append(oo); fh.append(len(A) != 0)
        fA.append(np.asarray(A) if len(A) else np.zeros((3, 3)))
    save("pose_find_transformation.npz", model=np.stack(fm), input=np.stack(fi), out=np.stack(fo), A=np.stack(fA),
         has_A=np.array(fh), npts=np.array([[5, 10, 7][k % 3] for k in range(200)]))
    # --- feature_scaling / handle_undetected_points
    fs = np.stack([ref_common.feature_scaling(ref_common.handle_undetected_points(p, p)[0]) for p in poses[:64]])
    save("pose_feature_scaling.npz", poses=poses[:64], scaled=fs)
    # --- procrustes (all option combinations) + superimpose / superimpose_old
    X, Y, D, Z, T = [], [], [], [], []
    good = poses[((poses != 0).any(2)).sum(1) >= 6]  # procrustes on fewer joints divides 0/0 in the reference
    opts = [(True, "best"), (False, "best"), (True, True), (True, False), (False, True), (False, False)]
    for k in range(120):
        a, b = good[rng.integers(len(good))], good[rng.integers(len(good))]
        sc, rf = opts[k % 6]
        d, z, tf = ref_procrustes.procrustes(a, b, sc, rf)  # procrustes.py:150
        X.append(a); Y.append(b); D.append(d); Z.append(z)
        T.append(np.concatenate([tf["rotation"].ravel(), [tf["scale"]], tf["translation"]]))
    so_Z, si_Z, si_n = [], [], []
    while len(so_Z) < 60:
        a, b = good[rng.integers(len(good))], good[rng.integers(len(good))]
        if (((a != 0).any(1)) & ((b != 0).any(1))).sum() < 3:
            continue
        X.append(a); Y.append(b)
        z, _ = ref_procrustes.superimpose_old(a, b)  # procrustes.py:68
        so_Z.append(z)
        z2, _ = ref_procrustes.superimpose(a, b)  # procrustes.py:7
        pad = np.full((18, 2), np.nan)
        pad[:len(z2)] = z2
        si_Z.append(pad); si_n.append(len(z2))
    save("pose_procrustes.npz", X=np.stack(X), Y=np.stack(Y), d=np.array(D), Z=np.stack(Z), tform=np.stack(T),
         scaling=np.array([o[0] for o in opts]), reflection=np.array([0 if o[1] == "best" else (1 if o[1] else 2) for o in opts]),
         superimpose_old_Z=np.stack(so_Z), superimpose_Z=np.stack(si_Z), superimpose_n=np.array(si_n))
    # --- pose_comparison decisions
    rows = []
    for k in range(200):
        A = np.eye(3) + 0.4 * rng.standard_normal((3, 3))
        d, sh = rng.uniform(0, 0.4), rng.uniform(0, 0.25)
        rows.append(np.concatenate([A.ravel(), [d, sh,
                                                ref_pc.decide_torso_shoulders_incl(d, A, 0.236, 19.8, sh, 0.125),
                                                ref_pc.decide_legs(d, A, 0.082, 24)]]))
    save("pose_decide.npz", rows=np.stack(rows))


def record_json_parser():
    """common.parse_JSON_multi_person (common.py:215-239) on a few fixture files: raw keypoint triples in, poses out."""
    raws, outs, counts = [], [], []
    for name in ("duo1.json", "trio1.json", "foto3.json", "1027.json", "emma1.json", "taj3.json"):
        f = os.path.join(REF, "json_data", name)
        if not os.path.exists(f):
            continue
        people = json.load(open(f))["people"]
        ref = ref_common.parse_JSON_multi_person(f)
        raw = np.zeros((8, 54)); out = np.zeros((8, 18, 2))
        n = min(len(people), 8)
        for k in range(n):
            raw[k] = people[k]["pose_keypoints"]
            out[k] = ref[k]
        raws.append(raw); outs.append(out); counts.append(n)
    save("pose_json_parser.npz", raw=np.stack(raws), poses=np.stack(outs), n=np.array(counts))


def record_scene_score():
    """urban_scene.match_scene_multi steps 3.1-5 (urban_scene.py:80-145): transformation.perspective_correction +
    transformation.affine_multi with random.sample pinned to recorded picks."""
    import random
    import urbanscene.transformation as ref_tr
    import plot_vars
    poses, _ = load_fixture_poses()
    full = poses[((poses != 0).any(2)).sum(1) == 18]
    rng = np.random.default_rng(50)
    out = {k: [] for k in ("src", "dst", "H2", "model_pose", "input_pose", "picks", "score", "n")}
    model_img = np.zeros((480, 640), np.uint8)
    for case in range(16):
        src, dst, Ht = synth.ransac_pair(120, inlier_ratio=1.0, seed=60 + case, noise=0.7)
        H2, _ = cv2.findHomography(dst.reshape(-1, 1, 2), src.reshape(-1, 1, 2), cv2.RANSAC, 5.0)
        model_pose = full[rng.integers(len(full))] * 0.5 + 20.0
        if case % 4 == 3:
            model_pose = model_pose.copy(); model_pose[[3, 9]] = 0       # undetected joints
        pp = np.c_[model_pose, np.ones(18)] @ Ht.T
        input_pose = pp[:, :2] / pp[:, 2:] + rng.normal(0, 1.5, (18, 2))
        input_pose[(model_pose == 0).all(1)] = 0
        p_model_good, p_input_good = src, dst
        picks = sorted(rng.choice(len(src), 5, replace=False).tolist())
        (p_trans, pose_trans, _img) = ref_tr.perspective_correction(H2, np.vstack((p_model_good, model_pose)),
                                                                    np.vstack((p_input_good, input_pose)), model_pose,
                                                                    input_pose, model_img, model_img, False)
        only_buildings = p_trans[0:len(p_trans) - 18]
        plot_vars.input_background_no_correction = p_input_good
        orig = random.sample
        random.sample = lambda population, k: list(picks)
        try:
            score = ref_tr.affine_multi(p_model_good, only_buildings, model_pose, pose_trans, 480, 640, 480, 640, "g",
                                        model_img, model_img, input_pose, False)
        finally:
            random.sample = orig
        for k, v in (("src", src), ("dst", dst), ("H2", H2), ("model_pose", model_pose), ("input_pose", input_pose),
                     ("picks", np.array(picks, np.int32)), ("score", score), ("n", len(src))):
            out[k].append(v)
    save("scene_score.npz", src=np.concatenate(out["src"]), dst=np.concatenate(out["dst"]), H2=np.stack(out["H2"]),
         model_pose=np.stack(out["model_pose"]), input_pose=np.stack(out["input_pose"]), picks=np.stack(out["picks"]),
         score=np.array(out["score"]), pair_offsets=np.concatenate([[0], np.cumsum(out["n"])]).astype(np.int64),
         model_wh=np.array([640.0, 480.0]))


def record_multi_person():
    """multi_person.find_ordered_matches (multi_person.py:169) on the multi-person fixtures."""
    import posematching.multi_person as ref_mp
    groups = []
    for f in sorted(glob.glob(os.path.join(REF, "json_data", "*.json"))):
        try:
            people = [p for p in ref_common.parse_JSON_multi_person(f) if p.shape == (18, 2)]
        except Exception:
            continue
        if 2 <= len(people) <= 8 and all(((p != 0).any(1)).sum() >= 8 for p in people):
            groups.append(np.stack(people))
    rng = np.random.default_rng(40)
    cases = []
    for k in range(40):
        a, b = groups[rng.integers(len(groups))], groups[rng.integers(len(groups))]
        if k % 3 == 0:
            b = a + rng.normal(0, 3.0, a.shape) * (a != 0)  # same scene, jittered: guaranteed candidates
        try:
            md, mo, ip = ref_mp.find_ordered_matches(list(a), list(b))
        except Exception:
            continue
        cases.append((a, b, md, mo, ip))
    P = 8
    A = np.zeros((len(cases), P, 18, 2)); B = np.zeros((len(cases), P, 18, 2))
    na = np.zeros(len(cases), np.int32); nb = np.zeros(len(cases), np.int32)
    found = np.zeros(len(cases), np.int8)
    table = -np.ones((len(cases), P, P), np.int32)   # table[c, model, :] = matching input indices, -1 padded
    n_model = np.zeros(len(cases), np.int32)
    for c, (a, b, md, mo, ip) in enumerate(cases):
        A[c, :len(a)], B[c, :len(b)], na[c], nb[c] = a, b, len(a), len(b)
        if md is not None:
            found[c], n_model[c] = 1, len(mo)
            for m, lst in md.items():
                table[c, m, :len(lst)] = lst
    save("pose_multi_person.npz", A=A, B=B, na=na, nb=nb, found=found, table=table, n_model=n_model)


def record_scene_score_single():
    """urban_scene.match_scene_single_person steps 4-5 (urban_scene.py:196-240): transformation.perspective_correction +
    transformation.affine_trans_interaction_pose_rand_scene (transformation.py:736-790: torso and legs each fitted
    together with background matches 0, 1 and 10; min/max-normalised residuals)."""
    import urbanscene.transformation as ref_tr
    poses, _ = load_fixture_poses()
    full = poses[((poses != 0).any(2)).sum(1) == 18]
    rng = np.random.default_rng(70)
    out = {k: [] for k in ("src", "dst", "H2", "model_pose", "input_pose", "result", "n")}
    img = np.zeros((480, 640), np.uint8)
    for case in range(20):
        src, dst, Ht = synth.ransac_pair(40 + 7 * case, inlier_ratio=1.0, seed=80 + case, noise=0.7)
        H2, _ = cv2.findHomography(dst.reshape(-1, 1, 2), src.reshape(-1, 1, 2), cv2.RANSAC, 5.0)
        model_pose = full[rng.integers(len(full))] * 0.5 + 20.0
        if case % 5 == 4:
            model_pose = model_pose.copy(); model_pose[[4, 12]] = 0      # undetected joints
        pp = np.c_[model_pose, np.ones(18)] @ Ht.T
        input_pose = pp[:, :2] / pp[:, 2:] + rng.normal(0, 1.0 + case % 3, (18, 2))
        input_pose[(model_pose == 0).all(1)] = 0
        (p_trans, pose_trans, _img) = ref_tr.perspective_correction(H2, np.vstack((src, model_pose)), np.vstack((dst, input_pose)),
                                                                    model_pose, input_pose, img, img, False)
        only_buildings = p_trans[0:len(p_trans) - 18]
        res = ref_tr.affine_trans_interaction_pose_rand_scene(src, only_buildings, model_pose, pose_trans, img, img, "g", False)
        for k, v in (("src", src), ("dst", dst), ("H2", H2), ("model_pose", model_pose), ("input_pose", input_pose),
                     ("result", np.array(res, np.float64)), ("n", len(src))):
            out[k].append(v)
    save("scene_score_single.npz", src=np.concatenate(out["src"]), dst=np.concatenate(out["dst"]), H2=np.stack(out["H2"]),
         model_pose=np.stack(out["model_pose"]), input_pose=np.stack(out["input_pose"]), result=np.stack(out["result"]),
         pair_offsets=np.concatenate([[0], np.cumsum(out["n"])]).astype(np.int64))


def record_multi_person_match():
    """multi_person.match (multi_person.py:15-166): the permutation stage behind find_ordered_matches -- Procrustes
    superimposition of every (model person, input person) of every injective permutation, min/max scaling of the stacked
    poses, one affine fit of the whole group, the MP_DISCTANCE gate -- on the multi-person fixtures."""
    import posematching.multi_person as ref_mp
    groups = []
    for f in sorted(glob.glob(os.path.join(REF, "json_data", "*.json"))):
        try:
            people = [p for p in ref_common.parse_JSON_multi_person(f) if p.shape == (18, 2)]
        except Exception:
            continue
        if 2 <= len(people) <= 5 and all(((p != 0).any(1)).sum() >= 8 for p in people):
            groups.append(np.stack(people))
    rng = np.random.default_rng(41)
    cases = []
    for k in range(60):
        a = groups[rng.integers(len(groups))]
        if k % 4 == 3:
            b = groups[rng.integers(len(groups))]          # unrelated group: mostly no candidates
        else:
            s, th = rng.uniform(0.8, 1.25), rng.uniform(-0.12, 0.12)
            R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            b = ((a @ R.T) * s + rng.uniform(-30, 30, 2) + rng.normal(0, 2.5, a.shape)) * (a != 0).any(2, keepdims=True)
            if k % 4 == 1 and len(b) > 2:
                b = b[rng.permutation(len(b))]             # people in a different order
            if k % 4 == 2:
                b = np.concatenate([b, groups[rng.integers(len(groups))][:1]])  # one extra (background) person
        try:
            md, mo, ip = ref_mp.find_ordered_matches(list(a), list(b))
            res = ref_mp.match(list(a), list(b))
        except Exception:
            continue
        cases.append((a, b, md, mo, ip, res))
    P, NP = 6, 16
    n = len(cases)
    A = np.zeros((n, P, 18, 2)); B = np.zeros((n, P, 18, 2))
    na = np.zeros(n, np.int32); nb = np.zeros(n, np.int32)
    found = np.zeros(n, np.int8); n_model = np.zeros(n, np.int32); n_input = np.zeros(n, np.int32)
    table = -np.ones((n, P, P), np.int32)
    MO = np.zeros((n, P, 18, 2)); IP = np.zeros((n, P, 18, 2))
    ok = np.zeros(n, np.int8); nperm = np.zeros(n, np.int32)
    perms = -np.ones((n, NP, P), np.int32); scores = np.full((n, NP), np.nan)
    for c, (a, b, md, mo, ip, res) in enumerate(cases):
        A[c, :len(a)], B[c, :len(b)], na[c], nb[c] = a, b, len(a), len(b)
        if md is not None:
            found[c], n_model[c], n_input[c] = 1, len(mo), len(ip)
            MO[c, :len(mo)], IP[c, :len(ip)] = np.stack(mo), np.stack(ip)
            for m, lst in md.items():
                table[c, m, :len(lst)] = lst
        ok[c] = 1 if res.match_bool else 0
        if res.matching_permutations:
            items = sorted(res.matching_permutations.items())
            assert len(items) <= NP
            nperm[c] = len(items)
            for k, (perm, val) in enumerate(items):
                perms[c, k, :len(perm)] = perm
                scores[c, k] = val["score"]
    print("multi_person.match: %d cases, %d matched, %d permutations kept" % (n, int(ok.sum()), int(nperm.sum())))
    save("pose_multi_person_match.npz", A=A, B=B, na=na, nb=nb, found=found, n_model=n_model, n_input=n_input, table=table,
         MO=MO, IP=IP, ok=ok, nperm=nperm, perms=perms, scores=scores, mp_distance=np.array(0.38))


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for fn in (lambda: record_scene_pair("pisa1_pisa2", "pisa1.jpg", "pisa2.jpg"), record_knn_synthetic, record_ransac_cv2,
               record_pose, record_multi_person, record_json_parser, record_scene_score, record_scene_pairs_sift_extra,
               record_scene_pair_orb_extra, record_multi_person_match, record_scene_score_single, record_ransac_real_pairs,
               record_ransac_synth_cv2, record_validate_homography, record_orb_pattern,
               record_orb_descriptors):
        name = getattr(fn, "__name__", "record_scene_pair")
        if not only or name in only or (name == "<lambda>" and "record_scene_pair" in only):
            fn()
