"""Candidate stage of posematching/multi_person.py, same signatures: order_poses (:225-253) and
find_ordered_matches (:169-207).  The reference calls match_single once per (model person, input person); here all
pairs go through ONE launch of hpusm_pose_match_pairs and the reference's loop is replayed over the result matrix.
multi_person.match itself (permutation search, :15-166) is control flow outside the replaced path."""
import numpy as np

from .. import thresholds
from .._dev import engine, up, down


def _leftmost_x(pose):
    xs = pose[:, 0]
    xs = xs[xs != 0]
    return xs.min() if xs.size else None


def order_poses(poses):
    """The reference's left-to-right ordering (multi_person.py:225-253), restated.  Poses with at most 8 non-zero
    coordinates are dropped.  Pose i walks back over its predecessors in the ORIGINAL list while it lies left of
    them; running off the front puts it first, otherwise it is inserted at position 1 (sic: the reference inserts at
    index 1, not behind the predecessor it stopped at); a predecessor or pose without any detected x ends the walk
    without inserting the pose at all."""
    ordered = []
    for i, pose in enumerate(poses):
        if np.count_nonzero(pose) <= 8:
            continue
        mine = _leftmost_x(pose)
        back = 1
        while True:
            j = i - back
            if j < 0:
                ordered.insert(0, pose)
                break
            theirs = _leftmost_x(poses[j])
            if mine is None or theirs is None:  # np.min of an empty selection raises in the reference: pose is skipped
                break
            if mine < theirs:
                back += 1
                continue
            ordered.insert(1, pose)
            break
    return ordered


def match_matrix(model_poses, input_poses):
    """(match bool [N, M], error_score float [N, M]) of every model person against every input person."""
    N, M = len(model_poses), len(input_poses)
    if N == 0 or M == 0:
        return np.zeros((N, M), bool), np.zeros((N, M))
    mo = np.repeat(np.asarray(model_poses, dtype=np.float64), M, axis=0)
    ip = np.tile(np.asarray(input_poses, dtype=np.float64), (N, 1, 1))
    thr = (thresholds.SP_DISTANCE_TORSO, thresholds.SP_ROTATION_TORSO, thresholds.SP_DISTANCE_LEGS,
           thresholds.SP_ROTATION_LEGS, thresholds.SP_DISTANCE_SHOULDER)
    match, score = engine.pose_match_pairs(up(mo, np.float64), up(ip, np.float64), thr)
    return down(match).astype(bool).reshape(N, M), down(score).astype(np.float64).reshape(N, M)


def find_ordered_matches(model_poses, input_poses):
    model_poses = order_poses(model_poses)
    input_poses = order_poses(input_poses)
    match, _ = match_matrix(model_poses, input_poses)
    matches = []
    matches_dict = {}
    for model_counter in range(0, len(model_poses)):
        matches.append([])
        start_input = 0
        if model_counter > 0:
            start_input = matches[model_counter - 1][0]
        model_matches = [j for j in range(start_input, len(input_poses)) if match[model_counter, j]]
        matches[model_counter].extend(model_matches)
        if not model_matches:
            return (None, None, None)
        matches_dict[model_counter] = model_matches
    return matches_dict, model_poses, input_poses
