"""NumPy restatement of the reference's pose-scoring arithmetic -- test infrastructure only.

Follows (file:line relative to /root/reference):
  common.find_transformation        common.py:37-101   (np.linalg.lstsq at :84)
  common.feature_scaling            common.py:688-708
  common.handle_undetected_points   common.py:281-348
  common.split_in_face_legs_torso_v2 / unsplit   common.py:269-279
  pose_comparison.*                 posematching/pose_comparison.py:8-81
  single_person.match_single        posematching/single_person.py:80-189
  procrustes.procrustes / superimpose / superimpose_old   posematching/procrustes.py:150-259, :7-54, :68-106
  multi_person.match (permutation stage)  posematching/multi_person.py:80-160

Pinned by tests/golden/pose_*.npz, which oracle/record_golden.py recorded from the UNMODIFIED reference
modules imported in the build container (numpy 2.3.5); tests/test_oracle_golden.py replays them.
"""
import math

import numpy as np

TORSO = [0, 1, 2, 3, 4, 5, 6, 7, 8, 11]
LEGS = [1, 8, 9, 10, 11, 12, 13]
FACE = [0, 14, 15, 16, 17]

# thresholds.py:7-13
SP_DISTANCE_TORSO, SP_ROTATION_TORSO = 0.236, 19.8
SP_DISTANCE_LEGS, SP_ROTATION_LEGS = 0.082, 24.0
SP_DISTANCE_SHOULDER = 0.125
DEFAULT_THRESHOLDS = (SP_DISTANCE_TORSO, SP_ROTATION_TORSO, SP_DISTANCE_LEGS, SP_ROTATION_LEGS, SP_DISTANCE_SHOULDER)


def find_transformation(model, inp):
    """(input_transform (n,2), A (3,3)) or (input, []) when no input row is detected."""
    model = np.asarray(model, dtype=np.float64)
    inp = np.asarray(inp, dtype=np.float64)
    detected = ~((inp[:, 0] == 0) & (inp[:, 1] == 0))
    if not detected.any():
        return inp, []
    X = np.hstack([inp[detected], np.ones((int(detected.sum()), 1))])
    Y = np.hstack([model[detected], np.ones((int(detected.sum()), 1))])
    A = np.linalg.lstsq(X, Y, rcond=None)[0]
    out = np.zeros_like(inp)
    out[detected] = (X @ A)[:, :2]
    A = A.copy()
    A[np.abs(A) < 1e-10] = 0
    return out, A


def feature_scaling(pose):
    """Min/max normalisation with the reference's quirk: each minimum runs over BOTH columns of the rows
    whose x (resp. y) is non-zero."""
    pose = np.asarray(pose, dtype=np.float64)
    xmax, ymax = pose[:, 0].max(), pose[:, 1].max()
    xmin = pose[pose[:, 0] != 0].min()
    ymin = pose[pose[:, 1] != 0].min()
    out = np.stack([(pose[:, 0] - xmin) / (xmax - xmin), (pose[:, 1] - ymin) / (ymax - ymin)], 1)
    out[out < 0] = 0
    return out


def handle_undetected_points(inp, model):
    """A joint that is (0,0) on either side becomes (0,0) on both; returns copies (input, model)."""
    inp = np.array(inp, dtype=np.float64)
    model = np.array(model, dtype=np.float64)
    und_in = (inp[:, 0] == 0) & (inp[:, 1] == 0)
    und_mo = (model[:, 0] == 0) & (model[:, 1] == 0)
    model[und_in] = 0
    inp[und_mo] = 0
    return inp, model


def split(pose):
    return pose[FACE], pose[TORSO], pose[LEGS]


def unsplit(face, torso, legs):
    return np.vstack([face[0], torso, legs, face[1:5]])


def _py_max(values):
    """Python's builtin max over a sequence (NaN semantics differ from np.max)."""
    it = iter(values)
    best = next(it)
    for v in it:
        if v > best:
            best = v
    return best


def euclidean_distances(model, transformed):
    d = np.abs(np.asarray(model) - np.asarray(transformed))
    return (d[:, 0] ** 2 + d[:, 1] ** 2) ** 0.5


def max_euclidean_distance(model, transformed):
    return _py_max(euclidean_distances(model, transformed))


def max_euclidean_distance_shoulders(model_torso, transformed_torso):
    e = euclidean_distances(model_torso, transformed_torso)
    return _py_max([e[1], e[4]])


def rotation_max(A):
    if len(A) == 0:
        return 0
    r1 = abs(math.atan2(-A[0][1], A[0][0]) * 57.3)
    r2 = abs(math.atan2(A[1][0], A[1][1]) * 57.3)
    return _py_max([r2, r1])


def decide_torso_shoulders_incl(d_torso, A, eucl_thr, rot_thr, d_shoulders, shoulder_thr):
    return bool(d_torso <= eucl_thr and rotation_max(A) <= rot_thr and d_shoulders <= shoulder_thr)


def decide_legs(err, A, eucl_thr, rot_thr):
    return bool(err <= eucl_thr and rotation_max(A) <= rot_thr)


def match_single(model, inp, thresholds=DEFAULT_THRESHOLDS):
    """Returns dict(match, error_score, input_transformation (22,2), plus the per-part intermediates)."""
    d_torso_thr, rot_torso_thr, d_legs_thr, rot_legs_thr, d_sh_thr = thresholds
    inp_c, model_c = handle_undetected_points(inp, model)
    model_s, inp_s = feature_scaling(model_c), feature_scaling(inp_c)
    m_face, m_torso, m_legs = split(model_s)
    i_face, i_torso, i_legs = split(inp_s)

    t_face, A_face = find_transformation(m_face, i_face)
    e_face = max_euclidean_distance(m_face, t_face)

    t_torso, A_torso = find_transformation(m_torso, i_torso)
    e_torso = max_euclidean_distance(m_torso, t_torso)
    e_sh = max_euclidean_distance_shoulders(m_torso, t_torso)
    if np.count_nonzero(m_torso) > 4:
        if np.count_nonzero(m_torso) - np.count_nonzero(i_torso) < 2:
            ok_torso = decide_torso_shoulders_incl(e_torso, A_torso, d_torso_thr, rot_torso_thr, e_sh, d_sh_thr)
        else:
            ok_torso = False
    else:
        ok_torso = True

    t_legs, A_legs = find_transformation(m_legs, i_legs)
    e_legs = max_euclidean_distance(m_legs, t_legs)
    if np.count_nonzero(m_legs) > 8:
        if np.count_nonzero(m_legs) - np.count_nonzero(i_legs) < 2:
            ok_legs = decide_legs(e_legs, A_legs, d_legs_thr, rot_legs_thr)
        else:
            ok_legs = False
    else:
        ok_legs = True

    return {
        "match": bool(ok_torso and ok_legs),
        "error_score": (e_torso + e_legs) / 2.0,
        "input_transformation": unsplit(t_face, t_torso, t_legs),
        "e_face": e_face, "e_torso": e_torso, "e_legs": e_legs, "e_shoulders": e_sh,
        "rot_torso": rotation_max(A_torso), "rot_legs": rotation_max(A_legs),
        "A_face": A_face, "A_torso": A_torso, "A_legs": A_legs,
    }


def match_single_batch(query, db, query_is_model=True, thresholds=DEFAULT_THRESHOLDS):
    """(match [n] bool, score [n] f64) of one query pose against db [n,18,2]."""
    n = len(db)
    match = np.zeros(n, bool)
    score = np.zeros(n, np.float64)
    for k in range(n):
        with np.errstate(all="ignore"):
            try:
                r = match_single(query, db[k], thresholds) if query_is_model else match_single(db[k], query, thresholds)
                match[k], score[k] = r["match"], r["error_score"]
            except ValueError:  # np.min of an empty selection: a pose without any detected joint
                match[k], score[k] = False, np.nan
    return match, score


def procrustes(X, Y, scaling=True, reflection="best"):
    """(d, Z, dict(rotation, scale, translation)) -- MATLAB-style Procrustes of Y onto X."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    muX, muY = X.mean(0), Y.mean(0)
    X0, Y0 = X - muX, Y - muY
    ssX, ssY = (X0 ** 2.0).sum(), (Y0 ** 2.0).sum()
    normX, normY = np.sqrt(ssX), np.sqrt(ssY)
    X0 = X0 / normX
    Y0 = Y0 / normY
    U, s, Vt = np.linalg.svd(X0.T @ Y0, full_matrices=False)
    V = Vt.T
    T = V @ U.T
    if not (isinstance(reflection, str) and reflection == "best"):
        if bool(reflection) != bool(np.linalg.det(T) < 0):
            V = V.copy()
            s = s.copy()
            V[:, -1] *= -1
            s[-1] *= -1
            T = V @ U.T
    trace = s.sum()
    if scaling:
        b = trace * normX / normY
        d = 1 - trace ** 2
        Z = normX * trace * (Y0 @ T) + muX
    else:
        b = 1
        d = 1 + ssY / ssX - 2 * trace * normY / normX
        Z = normY * (Y0 @ T) + muX
    c = muX - b * (muY @ T)
    return d, Z, {"rotation": T, "scale": b, "translation": c}


def superimpose(inp, model):
    """procrustes.superimpose: joints detected on both sides, no scaling, lowest-y rows aligned."""
    inp = np.asarray(inp, dtype=np.float64)
    model = np.asarray(model, dtype=np.float64)
    both = ~(((model[:18, 0] == 0) & (model[:18, 1] == 0)) | ((inp[:18, 0] == 0) & (inp[:18, 1] == 0)))
    model, inp = model[:18][both], inp[:18][both]
    _, Z, _ = procrustes(inp, model, False)
    Z[:, 1] += model[:, 1].min() - Z[:, 1].min()
    return Z, model


def superimpose_old(inp, model):
    """procrustes.superimpose_old: all rows (zeros included), then align the lower ankle in y."""
    inp = np.asarray(inp, dtype=np.float64)
    model = np.asarray(model, dtype=np.float64)
    _, Z, _ = procrustes(inp, model, False)
    foot = 10 if inp[10][1] >= inp[13][1] else 13
    Z[:, 1] += inp[foot][1] - Z[foot][1]
    return Z, model


def multi_person_permutations(model_poses, input_poses, matches_dict, mp_distance=0.38, normalise=True):
    """The permutation stage of multi_person.match (multi_person.py:80-160) on the ORDERED poses and candidate lists
    that find_ordered_matches returned: {permutation tuple: score} of the permutations that pass the MP_DISCTANCE gate."""
    import itertools
    out = {}
    for perm in itertools.product(*(matches_dict[k] for k in sorted(matches_dict))):
        if len(perm) != len(set(perm)):
            continue
        inputs, models = [], []
        for mi, ii in enumerate(perm):
            ip, mo = handle_undetected_points(input_poses[ii], model_poses[mi])
            Z, _ = superimpose_old(ip, mo)
            inputs.append(Z)
            models.append(mo)
        inputs, models = np.vstack(inputs), np.vstack(models)
        if normalise:
            inputs, models = feature_scaling(inputs), feature_scaling(models)
        full, _ = find_transformation(models, inputs)
        score = max_euclidean_distance(models, full)
        if score <= mp_distance:
            out[tuple(int(x) for x in perm)] = float(score)
    return out


def perspective_transform32(pts, H):
    """cv2.perspectiveTransform on float32 points with a float64 3x3 (features.py:77, transformation.py:38)."""
    p = np.asarray(pts, dtype=np.float32).astype(np.float64).reshape(-1, 2)
    w = p[:, 0] * H[2, 0] + p[:, 1] * H[2, 1] + H[2, 2]
    ok = np.abs(w) > np.finfo(np.float64).eps
    w = np.where(ok, 1.0 / np.where(ok, w, 1.0), 0.0)
    x = (p[:, 0] * H[0, 0] + p[:, 1] * H[0, 1] + H[0, 2]) * w
    y = (p[:, 0] * H[1, 0] + p[:, 1] * H[1, 1] + H[1, 2]) * w
    return np.stack([x, y], 1).astype(np.float32)


def scene_score(p_model_good, p_input_good, H2, model_pose, input_pose, model_w, model_h, picks):
    """urban_scene.match_scene_multi steps 3.1-5 (urban_scene.py:80-145) = transformation.perspective_correction
    (transformation.py:22-53) + transformation.affine_multi (:101-173,:413) for the given background picks."""
    p_model_good = np.asarray(p_model_good)
    if len(p_model_good) < 5:
        return 1
    bg_in = perspective_transform32(p_input_good, H2).astype(np.float64)
    pose_in = perspective_transform32(input_pose, H2).astype(np.float64)
    model_pts = np.vstack([np.asarray(model_pose, np.float64)] + [p_model_good[i][None].astype(np.float64) for i in picks])
    input_pts = np.vstack([pose_in] + [bg_in[i][None] for i in picks])
    _, M = find_transformation(model_pts, input_pts)
    trans = (np.hstack([input_pts, np.ones((len(input_pts), 1))]) @ M)[:, :2]
    a = np.stack([model_pts[:, 0] / model_w, model_pts[:, 1] / model_h], 1)
    b = np.stack([trans[:, 0] / model_w, trans[:, 1] / model_h], 1)
    return _py_max(euclidean_distances(a, b))


def split_v1(pose):
    """common.split_in_face_legs_torso (common.py:261-267): (face, torso = joints 1..7, legs = joints 8..13)."""
    pose = np.asarray(pose)
    return np.vstack([pose[0], pose[14:18]]), pose[1:8], pose[8:14]


def scene_score_single(p_model_good, p_input_good, H2, model_pose, input_pose):
    """urban_scene.match_scene_single_person steps 4-5 (urban_scene.py:196-240) = transformation.perspective_correction
    + transformation.affine_trans_interaction_pose_rand_scene (transformation.py:736-790):
    (max torso, max legs, sum torso, sum legs) of the min/max-normalised residuals."""
    p_model_good = np.asarray(p_model_good, np.float64)
    bg_in = perspective_transform32(p_input_good, H2).astype(np.float64)
    pose_in = perspective_transform32(input_pose, H2).astype(np.float64)
    _, m_torso, m_legs = split_v1(np.asarray(model_pose, np.float64))
    _, i_torso, i_legs = split_v1(pose_in)
    extra_m = np.vstack([p_model_good[0], p_model_good[1], p_model_good[10]])
    extra_i = np.vstack([bg_in[0], bg_in[1], bg_in[10]])
    out = []
    for m_part, i_part in ((m_torso, i_torso), (m_legs, i_legs)):
        m_part, i_part = np.vstack([m_part, extra_m]), np.vstack([i_part, extra_i])
        _, M = find_transformation(m_part, i_part)
        stacked = np.vstack([bg_in, i_part])
        trans = (np.hstack([stacked, np.ones((len(stacked), 1))]) @ M)[:, :2]
        err = euclidean_distances(feature_scaling(np.vstack([p_model_good, m_part])), feature_scaling(trans))
        out.append((_py_max(err), sum(err)))
    return out[0][0], out[1][0], out[0][1], out[1][1]
