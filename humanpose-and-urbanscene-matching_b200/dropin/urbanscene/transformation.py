"""urbanscene/transformation.py of the reference (hot subset), same signatures: perspective_correction (:22-99),
affine_multi (:101-173,:413) and affine_trans_interaction_pose_rand_scene (:736-790).  Plot branches and the image warp
used only for plots are not reproduced."""
import random

import numpy as np

from .. import common, thresholds
from .._dev import engine, up, down
from .features import euclidean_distance

# the reference keeps these in the plot_vars module global; affine_multi reads input_background_no_correction
input_background_no_correction = None


def perspective_correction(H2, p_model, p_input, model_pose_features, input_pose_features, model_img, input_img, plot=False):
    """Returns (p_input_persp_trans, input_pose_trans, perspective_transform_input); the third entry (the warped
    image, used by the reference for plots) is None here."""
    pts = np.float32(p_input).reshape(-1, 1, 2)
    out = down(engine.perspective_transform(up(pts, np.float32), up(np.asarray(H2, dtype=np.float64), np.float64)))
    p_input_persp_trans = np.squeeze(out[:])
    input_pose_trans = p_input_persp_trans[len(p_input_persp_trans) - len(input_pose_features):len(p_input_persp_trans)]
    return (p_input_persp_trans, input_pose_trans, None)


def affine_multi(p_model_good, p_input_good, model_pose, input_pose, model_image_height, model_image_width, input_image_h,
                 input_image_w, label, model_img, persp_input_img, input_pose_org, plot=False, rng=random):
    """Affine fit of the pose + AMOUNT_BACKGROUND_FEATURES random background matches; returns the maximum
    normalised distance (1 when there are too few matches).  `rng` (additive, defaults to the module the reference
    uses) makes the random background choice reproducible."""
    if len(p_model_good) < thresholds.AMOUNT_BACKGROUND_FEATURES:
        return 1
    picks = rng.sample(range(0, len(p_model_good)), thresholds.AMOUNT_BACKGROUND_FEATURES)
    model_rows = [np.array(model_pose)] + [np.array(p_model_good[i]) for i in picks]
    input_rows = [np.array(input_pose)] + [np.array(p_input_good[i]) for i in picks]
    model_pts = np.vstack(model_rows)
    input_pts = np.vstack(input_rows)
    (_, M) = common.find_transformation(model_pts, input_pts)
    pad = lambda x: np.hstack([x, np.ones((x.shape[0], 1))])  # noqa: E731
    input_trans = np.dot(pad(input_pts), M)[:, :-1]
    a = common.scene_feature_scaling(model_pts, model_image_width, model_image_height)
    b = common.scene_feature_scaling(input_trans, model_image_width, model_image_height)
    return max(euclidean_distance(a, b))


def affine_trans_interaction_pose_rand_scene(p_model_good, p_input_good, model_pose, input_pose, model_img, input_img, label,
                                             plot=False):
    """The single-person scene score (urban_scene.py:223): torso and legs are each fitted TOGETHER with background
    matches 0, 1 and 10; the fit is applied to all background matches + the part, both sides are min/max-normalised and
    compared.  Returns (max torso, max legs, sum torso, sum legs) of the normalised residuals.

    Two launches of hpusm_affine_lsq_batch (the parts differ in length) and two of hpusm_feature_scaling_batch (model
    side and transformed input side of a part share a launch)."""
    p_model_good = np.asarray(p_model_good, dtype=np.float64).reshape(-1, 2)
    p_input_good = np.asarray(p_input_good, dtype=np.float64).reshape(-1, 2)
    (_, model_torso, model_legs) = common.split_in_face_legs_torso(np.asarray(model_pose, dtype=np.float64))
    (_, input_torso, input_legs) = common.split_in_face_legs_torso(np.asarray(input_pose, dtype=np.float64))
    anchors = [0, 1, 10]
    result = []
    for model_part, input_part in ((model_torso, input_torso), (model_legs, input_legs)):
        model_part = np.vstack([model_part, p_model_good[anchors]])
        input_part = np.vstack([input_part, p_input_good[anchors]])
        (_, M) = common.find_transformation(model_part, input_part)
        everything = np.vstack([p_input_good, input_part])
        moved = np.dot(np.hstack([everything, np.ones((len(everything), 1))]), M)[:, :-1]
        both = np.stack([np.vstack([p_model_good, model_part]), moved])
        scaled = down(engine.feature_scaling_batch(up(both, np.float64)))
        err = euclidean_distance(scaled[0], scaled[1])
        result.append((max(err), sum(err)))
    return (result[0][0], result[1][0], result[0][1], result[1][1])
