"""Hot subset of the reference's common.py, same signatures.

find_transformation (common.py:37-101), feature_scaling (:688-708), scene_feature_scaling (:674-686),
handle_undetected_points (:281-348), split_in_face_legs_torso_v2 / unsplit (:269-279), anorm helpers (:107-110).
"""
import numpy as np

from ._dev import engine, up, down


def anorm2(a):
    return (a * a).sum(-1)


def anorm(a):
    return np.sqrt(anorm2(a))


def find_transformation(model_features, input_features):
    """Affine least squares  [input|1] * A = [model|1]  over the rows whose input point is not (0,0).
    Returns (input_transform (n,2) float64, A (3,3) float64), or (input_features, []) when nothing is detected."""
    model = np.asarray(model_features, dtype=np.float64).reshape(-1, 2)
    inp = np.asarray(input_features, dtype=np.float64).reshape(-1, 2)
    out, A, has = engine.affine_lsq_batch(up(model[None], np.float64), up(inp[None], np.float64))
    if not bool(down(has)[0]):
        return (input_features, [])
    return (down(out)[0], down(A)[0])


def feature_scaling(input):
    pts = np.asarray(input, dtype=np.float64).reshape(-1, 2)
    return down(engine.feature_scaling_batch(up(pts[None], np.float64)))[0]


def scene_feature_scaling(input, xmax, ymax):
    input = np.asarray(input, dtype=np.float64)
    return np.vstack([input[:, 0] / xmax, input[:, 1] / ymax]).T


def handle_undetected_points(input_features, model_features):
    """Copies; a joint that is (0,0) on either side becomes (0,0) on both (pure masking, no arithmetic)."""
    input_copy = np.array(input_features, copy=True)
    model_copy = np.array(model_features, copy=True)
    und_in = (np.asarray(input_features)[:, 0] == 0) & (np.asarray(input_features)[:, 1] == 0)
    und_mo = (np.asarray(model_features)[:, 0] == 0) & (np.asarray(model_features)[:, 1] == 0)
    model_copy[und_in] = 0
    input_copy[und_mo] = 0
    assert len(model_copy) == len(input_copy)
    return (input_copy, model_copy)


def split_in_face_legs_torso(features):
    """common.py:261-267: torso = neck .. left wrist (joints 1-7), legs = joints 8-13, face = nose + eyes/ears."""
    features = np.asarray(features)
    return (np.vstack([features[0], features[14:18]]), features[1:8], features[8:14])


def split_in_face_legs_torso_v2(features):
    torso = features[[0, 1, 2, 3, 4, 5, 6, 7, 8, 11]]
    legs = features[[1, 8, 9, 10, 11, 12, 13]]
    face = np.vstack([features[0], features[14:18]])
    return (face, torso, legs)


def unsplit(face, torso, legs):
    return np.vstack([face[0], torso, legs, face[1:5]])
