"""posematching/single_person.py of the reference: match_single (:80-189) with the same signature and result type.
The whole decision (masking, normalisation, three affine fits, distances, rotations, thresholds) is ONE launch of
hpusm_pose_match_batch_f64 with a database of one pose; match_single_batch is the additive batched entry."""
import collections

import numpy as np

from .. import thresholds
from .._dev import engine, up, down

MatchResult = collections.namedtuple("MatchResult", ["match_bool", "error_score", "input_transformation"])


def _thresholds():
    return (thresholds.SP_DISTANCE_TORSO, thresholds.SP_ROTATION_TORSO, thresholds.SP_DISTANCE_LEGS,
            thresholds.SP_ROTATION_LEGS, thresholds.SP_DISTANCE_SHOULDER)


def match_single(model_features, input_features, normalise=True):
    if not normalise:  # the reference returns before comparing (single_person.py:107-111)
        return MatchResult(None, error_score=0, input_transformation=None)
    model = np.asarray(model_features, dtype=np.float64)
    inp = np.asarray(input_features, dtype=np.float64)
    match, score, detail = engine.pose_match_batch(up(model, np.float64), up(inp[None], np.float64), True, _thresholds(),
                                                   want_detail=True)
    detail = down(detail)[0]
    if not np.isfinite(detail[7]) and not bool(down(match)[0]):
        # no detected joint / degenerate bounding box: the reference raises from np.min / LAPACK at this point
        raise ValueError("pose cannot be normalised (no detected joints or zero extent)")
    return MatchResult(bool(down(match)[0]), error_score=float(detail[7]), input_transformation=detail[35:].reshape(22, 2))


def match_single_batch(model_features, input_database):
    """One model pose against [n,18,2] input poses: (match bool [n], error_score float [n])."""
    db = np.asarray(input_database)
    dt = np.float64 if db.dtype == np.float64 else np.float32
    match, score = engine.pose_match_batch(up(np.asarray(model_features), dt), up(db, dt), True, _thresholds())
    return down(match).astype(bool), down(score)
