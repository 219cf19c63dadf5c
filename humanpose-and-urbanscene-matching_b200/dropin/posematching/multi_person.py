"""posematching/multi_person.py of the reference, same signatures: order_poses (:225-253), find_ordered_matches
(:169-207) and match (:15-166).

Candidate stage: the reference calls match_single once per (model person, input person); here all pairs go through ONE
launch of hpusm_pose_match_pairs and the reference's loop is replayed over the result matrix.

Permutation stage (match): the reference superimposes every (model person, input person) of every injective permutation
with Procrustes, stacks the poses, min/max-scales them, fits one affine map to the whole group and gates the largest
residual.  Here ALL permutations go through one launch each of hpusm_procrustes_batch (every person pair of every
permutation), hpusm_feature_scaling_batch (both stacked groups) and hpusm_affine_lsq_batch (one fit per permutation);
the enumeration of the permutations stays on the host."""
import collections
import itertools
import logging

import numpy as np

from .. import thresholds
from .. import common
from .._dev import engine, up, down
from .single_person import match_single

logger = logging.getLogger("multi_person")
MatchResultMulti = collections.namedtuple("MatchResult", ["match_bool", "error_score", "input_transformation",
                                                          "matching_permutations"])


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


def input_pose_solo_multiple_times(matches_dict, input_length):
    """multi_person.py:213-222, quirk included: np.argmax over the boolean vector is the INDEX of the first input pose
    that is the only candidate of more than one model pose, so the answer is False when that pose is input 0.  (The
    reference computes it and ignores it: the early return behind it is commented out.)"""
    solo = np.zeros(input_length)
    for matching_inputs in matches_dict.values():
        if len(matching_inputs) == 1:
            solo[matching_inputs[0]] += 1
    return bool(np.argmax(solo > 1))


def permutation_scores(model_poses, input_poses, permutations, normalise=True):
    """Score of every permutation (tuple: input index per model pose) in three launches: (scores [P] float64,
    unchanged_model [P, 18 K, 2], unchanged_input [P, 18 K, 2])."""
    P, K = len(permutations), len(permutations[0])
    inp = np.empty((P, K, 18, 2))
    mod = np.empty((P, K, 18, 2))
    for p, perm in enumerate(permutations):
        for mi, ii in enumerate(perm):
            inp[p, mi], mod[p, mi] = common.handle_undetected_points(input_poses[ii], model_poses[mi])
    # procrustes.superimpose_old (procrustes.py:68-106): the model superimposed on the input without scaling, then the
    # lower foot of the input and of the result aligned in y
    _, Z, _ = engine.procrustes_batch(up(inp.reshape(P * K, 18, 2), np.float64), up(mod.reshape(P * K, 18, 2), np.float64), False)
    Z = down(Z).reshape(P, K, 18, 2)
    foot = np.where(inp[:, :, 10, 1] >= inp[:, :, 13, 1], 10, 13)
    take = lambda a: np.take_along_axis(a[..., 1], foot[..., None], axis=2)[..., 0]  # y of the chosen foot, [P, K]
    Z[..., 1] += (take(inp) - take(Z))[..., None]
    inputs, models = Z.reshape(P, K * 18, 2), mod.reshape(P, K * 18, 2)
    if normalise:
        both = down(engine.feature_scaling_batch(up(np.concatenate([inputs, models]), np.float64)))
        inputs_s, models_s = both[:P], both[P:]
    else:
        inputs_s, models_s = inputs, models
    full, _, _ = engine.affine_lsq_batch(up(models_s, np.float64), up(inputs_s, np.float64))
    diff = np.abs(models_s - down(full))
    dist = (diff[..., 0] ** 2 + diff[..., 1] ** 2) ** 0.5
    # pose_comparison.max_euclidean_distance is Python's max(): NaN when the first distance is NaN, otherwise NaNs lose
    with np.errstate(all="ignore"):
        scores = np.where(np.isnan(dist[:, 0]), np.nan, np.nanmax(np.where(np.isnan(dist), -np.inf, dist), axis=1))
    return scores, models, inp.reshape(P, K * 18, 2)


def match(model_poses, input_poses, plot=False, input_image=None, model_image=None, normalise=True):
    model_poses = np.copy(model_poses)
    input_poses = np.copy(input_poses)
    fail = MatchResultMulti(False, error_score=0, input_transformation=None, matching_permutations=None)
    if len(input_poses) == 0 or len(model_poses) == 0:
        return fail
    if len(input_poses) < len(model_poses):
        return fail
    if len(input_poses) > len(model_poses):
        logger.warning(" !! WARNING !! Amount of input poses > model poses")
    if len(model_poses) == 1 and len(input_poses) == 1:
        single = match_single(model_poses[0], input_poses[0], True)
        (model_pose, input_pose) = common.handle_undetected_points(model_poses[0], input_poses[0])
        if single.match_bool:
            return MatchResultMulti(single.match_bool, error_score=single.error_score,
                                    input_transformation=single.input_transformation,
                                    matching_permutations={0: {"score": single.error_score, "model": model_pose,
                                                               "input": input_pose}})
        return fail
    matches_dict, model_poses, input_poses = find_ordered_matches(list(model_poses), list(input_poses))
    if matches_dict is None:
        return fail
    combinations = list(itertools.product(*(matches_dict[name] for name in sorted(matches_dict))))
    injective = [perm for perm in combinations if len(perm) == len(set(perm))]
    result_permutations = {}
    if injective:
        scores, unchanged_model, unchanged_input = permutation_scores(model_poses, input_poses, injective, normalise)
        for p, perm in enumerate(injective):
            if scores[p] <= thresholds.MP_DISCTANCE:
                result_permutations[perm] = {"score": float(scores[p]), "model": unchanged_model[p], "input": unchanged_input[p]}
    return MatchResultMulti(bool(result_permutations), error_score=0, input_transformation=None,
                            matching_permutations=result_permutations)
