"""urbanscene/transformation.py of the reference (hot subset), same signatures: perspective_correction (:22-99) and
affine_multi (:101-173,:413).  Plot branches and the image warp used only for plots are not reproduced."""
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
