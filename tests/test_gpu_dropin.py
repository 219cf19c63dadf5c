"""The drop-in modules (reference names and signatures) against the golden vectors recorded from the reference."""
import importlib

import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import knn_oracle, ransac_oracle

pytestmark = pytest.mark.gpu
DROPIN = "humanpose-and-urbanscene-matching_b200.dropin"


@pytest.fixture(scope="module")
def mods(cuda_device):
    names = ["common", "thresholds", "urbanscene.features", "urbanscene.transformation", "posematching.procrustes",
             "posematching.pose_comparison", "posematching.single_person"]
    return {n: importlib.import_module(DROPIN + "." + n) for n in names}


def test_match_single_equals_reference_float64(mods):
    """single_person.match_single(model, input) on the reference's own float64 poses: same decision, same score,
    same (22,2) transformed input."""
    sp = mods["posematching.single_person"]
    g = load_golden("pose_match_single.npz")
    n = 0
    for k in range(0, len(g["model"]), 2):
        ref = int(g["match"][k])
        if ref == -1:
            with pytest.raises(ValueError):
                sp.match_single(g["model"][k], g["input"][k])
            continue
        if not np.isfinite(g["score"][k]):
            continue
        r = sp.match_single(g["model"][k], g["input"][k])
        assert r.match_bool == bool(ref), k
        np.testing.assert_allclose(r.error_score, g["score"][k], rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(r.input_transformation, g["input_transformation"][k], rtol=1e-6, atol=1e-8)
        n += 1
    assert n > 300
    assert sp.match_single(g["model"][0], g["input"][0], normalise=False).match_bool is None


def test_common_functions_equal_reference(mods):
    common = mods["common"]
    g = load_golden("pose_find_transformation.npz")
    for k in range(0, len(g["npts"]), 3):
        n = int(g["npts"][k])
        out, A = common.find_transformation(g["model"][k][:n], g["input"][k][:n])
        assert (len(A) != 0) == bool(g["has_A"][k])
        if len(A):
            np.testing.assert_allclose(out, g["out"][k][:n], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(A, g["A"][k], rtol=1e-4, atol=1e-7)
    g = load_golden("pose_feature_scaling.npz")
    for p, s in list(zip(g["poses"], g["scaled"]))[:20]:
        got = common.feature_scaling(common.handle_undetected_points(p, p)[0])
        np.testing.assert_allclose(got, s, rtol=1e-12, equal_nan=True)
    face, torso, legs = common.split_in_face_legs_torso_v2(g["poses"][0])
    assert common.unsplit(face, torso, legs).shape == (22, 2)


def test_procrustes_equals_reference(mods):
    pr = mods["posematching.procrustes"]
    g = load_golden("pose_procrustes.npz")
    for k in range(0, len(g["d"]), 5):
        rf = int(g["reflection"][k % 6])
        d, Z, tf = pr.procrustes(g["X"][k], g["Y"][k], bool(g["scaling"][k % 6]), "best" if rf == 0 else (rf == 1))
        np.testing.assert_allclose(d, g["d"][k], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(Z, g["Z"][k], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(tf["rotation"].ravel(), g["tform"][k][:4], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(tf["translation"], g["tform"][k][5:], rtol=1e-4, atol=1e-6)
    n0 = len(g["d"])
    for j in range(0, len(g["superimpose_old_Z"]), 4):
        a, b = g["X"][n0 + j], g["Y"][n0 + j]
        Z, _ = pr.superimpose_old(a, b)
        np.testing.assert_allclose(Z, g["superimpose_old_Z"][j], rtol=1e-4, atol=1e-6)
        Z2, _ = pr.superimpose(a, b)
        np.testing.assert_allclose(Z2, g["superimpose_Z"][j][:len(Z2)], rtol=1e-4, atol=1e-6)


def test_pose_comparison_equals_reference(mods):
    pc = mods["posematching.pose_comparison"]
    for r in load_golden("pose_decide.npz")["rows"]:
        A, d, sh = r[:9].reshape(3, 3), r[9], r[10]
        assert pc.decide_torso_shoulders_incl(d, A, 0.236, 19.8, sh, 0.125) == bool(r[11])
        assert pc.decide_legs(d, A, 0.082, 24) == bool(r[12])
    assert pc.decide_legs(0.01, [], 0.082, 24) is True


@pytest.mark.parametrize("det", ["sift", "orb"])
def test_match_and_draw_scene_pair(mods, det):
    """C1 through features.match_and_draw: kNN + ratio are pinned on the reference (cv2) bit for bit; the homography
    stage is pinned on the oracle under the shared sample seed (cv2's sampler is private)."""
    ft = mods["urbanscene.features"]
    g = load_golden("scene_pisa1_pisa2_%s.npz" % det)
    matcher = ft.BFMatcher(ft.NORM_L2 if det == "sift" else ft.NORM_HAMMING)
    dm = g["desc_model"].astype(np.float32) if det == "sift" else g["desc_model"]
    di = g["desc_input"].astype(np.float32) if det == "sift" else g["desc_input"]
    raw = matcher.knnMatch(dm, trainDescriptors=di, k=2)
    assert [m[0].trainIdx for m in raw] == g["knn_idx"][:, 0].tolist()
    assert np.array_equal(np.float32([m[1].distance for m in raw]), g["knn_dist"][:, 1])
    p1, p2, pairs = ft.filter_matches(g["kp_model"], g["kp_input"], raw)
    np.testing.assert_array_equal(p1, g["p_model"])
    np.testing.assert_array_equal(p2, g["p_input"])
    model_img = np.zeros(tuple(g["model_shape"]), np.uint8)
    input_img = np.zeros(tuple(g["input_shape"]), np.uint8)
    mask, good_m, good_i, H, H2, n = ft.match_and_draw("w", matcher, dm, di, g["kp_model"], g["kp_input"], model_img,
                                                       input_img, False)
    src, dst = g["p_model"].reshape(-1, 2), g["p_input"].reshape(-1, 2)
    # the drop-in replays OpenCV's sample sequence: the mask is the one the REFERENCE's match_and_draw returned; this
    # pair's fits are ill posed (validate_homography False, LM not converged), so H is held to the oracle's, loosely
    np.testing.assert_array_equal(mask, g["mask"])
    rH, rmask, rn, _ = ransac_oracle.homography_batch(src, dst, [0, len(src)], 5.0, 2000, 0.995, 0, True, 1)
    rH2, _, _, _ = ransac_oracle.homography_batch(dst, src, [0, len(src)], 5.0, 2000, 0.995, 0, True, 1)
    np.testing.assert_array_equal(mask.ravel(), rmask)
    np.testing.assert_allclose(H, rH[0], rtol=1e-2, atol=1e-2 * np.abs(rH[0]).max())
    np.testing.assert_allclose(H2, rH2[0], rtol=1e-2, atol=1e-2 * np.abs(rH2[0]).max())
    assert n == int(rn[0]) == int(g["n_good"]) == len(good_m) == len(good_i) and mask.shape == (len(src), 1) and mask.dtype == np.uint8
    np.testing.assert_array_equal(good_m, g["good_model_pts"])
    np.testing.assert_array_equal(good_i, g["good_input_pts"])
    assert input_img.any()  # the projected model outline was drawn into input_img, as the reference does
    assert ft.validate_homography(g["H"]) == bool(g["valid"])
    # too few matches -> the reference's sentinel tuple
    out = ft.match_and_draw("w", matcher, dm[:20], di, g["kp_model"][:20], g["kp_input"], model_img, input_img, False)
    assert out == (None, None, None, None, None, None)


def test_bfmatcher_akaze_width_and_flann_names(mods):
    """AKAZE's 61-byte binary descriptors through the BFMatcher duck-type (zero-padded to the kernel's 64-byte rows), and
    the '-flann' matcher names the reference's driver scripts use (urbanscene/main.py:15) resolve to the exact matcher."""
    ft = mods["urbanscene.features"]
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (300, 61), dtype=np.uint8)
    t = rng.integers(0, 256, (500, 61), dtype=np.uint8)
    t[:100] = q[:100] ^ (rng.random((100, 61)) < 0.02).astype(np.uint8)
    raw = ft.BFMatcher(ft.NORM_HAMMING).knnMatch(q, trainDescriptors=t, k=2)
    ri, rd = knn_oracle.knn2_hamming(q, t)
    assert [[m.trainIdx for m in r] for r in raw] == ri.tolist()
    assert [[int(m.distance) for m in r] for r in raw] == rd.tolist()
    det, matcher = ft.init_feature("orb-flann")
    assert isinstance(matcher, ft.BFMatcher) and matcher.normType == ft.NORM_HAMMING


def test_orb_detector_dropin_equals_cv2(mods):
    """init_feature('orb-flann') (the name the reference's drivers pass) hands out a detector whose detectAndCompute gives
    cv2's keypoints and, from the GPU, cv2's descriptors bit for bit (urban_scene.py:31-32)."""
    cv2 = pytest.importorskip("cv2")
    ft = mods["urbanscene.features"]
    detector, matcher = ft.init_feature("orb-flann")
    assert isinstance(detector, ft.OrbGpu)
    ref = cv2.ORB_create(nfeatures=100000, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_FAST_SCORE, patchSize=31)
    g = load_golden("orb_descriptors.npz")
    for k in (0, 1):
        img = g["img_%d" % k]
        kp, desc = detector.detectAndCompute(img, None)
        kp_ref, desc_ref = ref.detectAndCompute(img, None)
        assert isinstance(kp, ft.KeyPointArray) and len(kp) == len(kp_ref)
        assert [(q.pt, q.octave, q.angle, q.response, q.size) for q in kp] == \
               [(q.pt, q.octave, q.angle, q.response, q.size) for q in kp_ref]
        assert [(q.pt, q.angle) for q in kp.to_cv2()[:5]] == [(q.pt, q.angle) for q in kp_ref[:5]]
        kd = detector.detect(img, None)
        assert [q.pt for q in kd] == [q.pt for q in kp_ref]
        np.testing.assert_array_equal(desc, desc_ref)
        np.testing.assert_array_equal(desc, g["desc_%d" % k])
    assert detector.getNLevels() == 8  # attribute access falls through to the OpenCV detector
    # no silent CPU path: a detection mask (the reference never passes one) raises unless the caller opts in explicitly
    mask = np.full(img.shape, 255, np.uint8)
    with pytest.raises(NotImplementedError):
        detector.detectAndCompute(img, mask)
    os.environ["HPUSM_ORB_CPU_DETECT"] = "1"
    try:
        kp_m, desc_m = detector.detectAndCompute(img, mask)   # OpenCV detects, the GPU describes: still cv2's answer
        np.testing.assert_array_equal(desc_m, desc_ref)
    finally:
        del os.environ["HPUSM_ORB_CPU_DETECT"]


def test_transformation_functions(mods):
    tr = mods["urbanscene.transformation"]
    g = load_golden("perspective_transform.npz")
    pts = g["pts"].reshape(-1, 2)
    p_trans, pose_trans, _ = tr.perspective_correction(g["H"], pts, pts, pts[:18], pts[-18:], None, None)
    np.testing.assert_allclose(p_trans, g["out"].reshape(-1, 2), rtol=1e-6, atol=1e-5)
    np.testing.assert_array_equal(pose_trans, p_trans[-18:])
    assert tr.affine_multi(pts[:3], pts[:3], pts[:18], pts[:18], 480, 640, 480, 640, "x", None, None, None) == 1
    import random
    A = np.array([[0.9, 0.1], [-0.1, 0.9]])
    model_bg, model_pose = pts[:40].astype(np.float64), pts[40:58].astype(np.float64)
    input_bg, input_pose = model_bg @ A + 5.0, model_pose @ A + 5.0
    s = tr.affine_multi(model_bg, input_bg, model_pose, input_pose, 480, 640, 480, 640, "x", None, None, None,
                        rng=random.Random(3))
    assert s < 1e-9  # an exact affine relation is recovered exactly


def test_multi_person_candidate_stage(cuda_device):
    """multi_person.find_ordered_matches (multi_person.py:169-207): same candidate lists as the reference, from ONE
    launch over all (model person, input person) pairs."""
    mp = importlib.import_module(DROPIN + ".posematching.multi_person")
    g = load_golden("pose_multi_person.npz")
    n_found = 0
    for c in range(len(g["found"])):
        a = [p for p in g["A"][c][:g["na"][c]]]
        b = [p for p in g["B"][c][:g["nb"][c]]]
        md, mo, ip = mp.find_ordered_matches(a, b)
        if not g["found"][c]:
            assert md is None
            continue
        n_found += 1
        assert md is not None and len(mo) == g["n_model"][c]
        for m in range(len(mo)):
            ref = [int(x) for x in g["table"][c, m] if x >= 0]
            assert md[m] == ref, (c, m)
    assert n_found >= 10


def test_scene_score_single_person(cuda_device):
    """urban_scene.match_scene_single_person steps 4-5 through the drop-in transformation module: perspective_correction +
    affine_trans_interaction_pose_rand_scene against the reference's own results on 20 scenes."""
    tr = importlib.import_module(DROPIN + ".urbanscene.transformation")
    g = load_golden("scene_score_single.npz")
    offs = g["pair_offsets"]
    for c in range(len(g["result"])):
        a, b = int(offs[c]), int(offs[c + 1])
        src, dst, mp_, ip_ = g["src"][a:b], g["dst"][a:b], g["model_pose"][c], g["input_pose"][c]
        (p_trans, pose_trans, _img) = tr.perspective_correction(g["H2"][c], np.vstack((src, mp_)), np.vstack((dst, ip_)), mp_, ip_,
                                                                None, None, False)
        got = tr.affine_trans_interaction_pose_rand_scene(src, p_trans[0:len(p_trans) - 18], mp_, pose_trans, None, None, "g")
        np.testing.assert_allclose(got, g["result"][c], rtol=1e-7, atol=1e-10)


def test_multi_person_match_permutation_stage(cuda_device):
    """multi_person.match (multi_person.py:15-166) end to end -- candidate stage, every injective permutation superimposed,
    scaled and fitted in three batched launches -- against results recorded from the reference: same verdict, same
    surviving permutations, same scores, same (model, input) arrays."""
    mp = importlib.import_module(DROPIN + ".posematching.multi_person")
    g = load_golden("pose_multi_person_match.npz")
    n_perm = 0
    for c in range(len(g["ok"])):
        a = [p for p in g["A"][c][:g["na"][c]]]
        b = [p for p in g["B"][c][:g["nb"][c]]]
        res = mp.match(a, b)
        assert bool(res.match_bool) == bool(g["ok"][c]), c
        want = {tuple(int(x) for x in g["perms"][c, k] if x >= 0): float(g["scores"][c, k]) for k in range(int(g["nperm"][c]))}
        got = res.matching_permutations or {}
        assert set(got) == set(want), (c, sorted(got), sorted(want))
        for perm, score in want.items():
            assert abs(got[perm]["score"] - score) <= 1e-9 * max(1.0, abs(score)), (c, perm)
            assert got[perm]["model"].shape == got[perm]["input"].shape == (18 * len(perm), 2)
            n_perm += 1
    assert n_perm >= 30
    # the 1 x 1 shortcut goes through match_single
    one = mp.match([g["A"][0][0]], [g["A"][0][0]])
    assert one.match_bool and list(one.matching_permutations) == [0]
    assert not mp.match([], [g["A"][0][0]]).match_bool and not mp.match(list(g["A"][0][:2]), [g["A"][0][0]]).match_bool


def test_pose_match_pairs_equals_batch(hpusm, cuda_device):
    import torch
    g = load_golden("pose_match_single.npz")
    M, I = torch.from_numpy(g["model"]).to(cuda_device), torch.from_numpy(g["input"]).to(cuda_device)
    match, score = hpusm.pose_match_pairs(M, I)
    match, score = match.cpu().numpy(), score.cpu().numpy()
    ok = (g["match"] >= 0) & np.isfinite(g["score"])
    np.testing.assert_array_equal(match[ok].astype(bool), g["match"][ok].astype(bool))
    np.testing.assert_allclose(score[ok], g["score"][ok], rtol=1e-6, atol=1e-7)
    assert (match[g["match"] < 0] == 0).all()
