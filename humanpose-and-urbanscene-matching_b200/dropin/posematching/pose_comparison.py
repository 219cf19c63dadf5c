"""posematching/pose_comparison.py of the reference (:8-81), same signatures.  These are scalar decisions on values
the kernels produced; the batched path evaluates the same expressions inside pose_match_batch."""
import math

import numpy as np


def _rot_max(transformation_matrix):
    if not len(transformation_matrix) == 0:
        rotation_1 = np.abs(math.atan2(-transformation_matrix[0][1], transformation_matrix[0][0]) * 57.3)
        rotation_2 = np.abs(math.atan2(transformation_matrix[1][0], transformation_matrix[1][1]) * 57.3)
        return max(rotation_2, rotation_1)
    return 0


def decide_torso_shoulders_incl(max_euclid_distance_torso, transformation_matrix, eucld_tresh, rotation_tresh,
                                max_euclid_distance_shoulders, shoulder_thresh):
    rot_max = _rot_max(transformation_matrix)
    if max_euclid_distance_torso <= eucld_tresh and rot_max <= rotation_tresh:
        if max_euclid_distance_shoulders <= shoulder_thresh:
            return True
    return False


def decide_legs(max_error, transformation_matrix, eucld_tresh, rotation_tresh):
    rot_max = _rot_max(transformation_matrix)
    if max_error <= eucld_tresh and rot_max <= rotation_tresh:
        return True
    return False


def max_euclidean_distance_shoulders(model_torso, input_transformed_torso):
    e = np.abs(model_torso - input_transformed_torso)
    d = ((e[:, 0]) ** 2 + e[:, 1] ** 2) ** 0.5
    return max([d[1], d[4]])


def max_euclidean_distance(model, transformed_input):
    e = np.abs(model - transformed_input)
    d = ((e[:, 0]) ** 2 + e[:, 1] ** 2) ** 0.5
    return max(d)
