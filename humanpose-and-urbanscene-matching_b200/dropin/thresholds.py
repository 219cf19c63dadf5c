"""The reference's configuration system is one module of constants (thresholds.py:3-32); callers read them as module
attributes and the dataset scripts overwrite them at run time, so they are plain module globals here too.  Values are
the reference's, keyed by the line they come from."""

_REFERENCE_VALUES = (
    # name, value, thresholds.py line
    ("OPENPOSE_ZEKERHEID", 0.2, 3),            # minimum OpenPose joint confidence
    ("OPENPOSE_AMOUNT_KEYPOINTS", 2, 4),
    ("SP_DISTANCE_TORSO", 0.236, 7),           # single-person: max normalised distance / rotation (degrees*57.3/180..)
    ("SP_ROTATION_TORSO", 19.8, 8),
    ("SP_DISTANCE_LEGS", 0.082, 10),
    ("SP_ROTATION_LEGS", 24, 11),
    ("SP_DISTANCE_SHOULDER", 0.125, 13),
    ("MP_DISCTANCE", 0.38, 18),                # multi-person global affine fit
    ("MIN_MATCH_COUNT", 16, 21),               # ratio survivors needed before RANSAC
    ("FLANN_INDEX_KDTREE", 1, 22),
    ("FLANN_INDEX_LSH", 6, 23),
    ("FILTER_RATIO", 0.8, 24),                 # Lowe ratio
    ("PERSPECTIVE_CORRECTION", True, 28),
    ("USI_AMOUNT_ITERATIONS", 1, 29),
    ("AFFINE_TRANS_WHOLE_DISTANCE", 0.084, 30),
    ("AMOUNT_BACKGROUND_FEATURES", 5, 31),
    ("CROP", False, 32),
)

globals().update({name: value for name, value, _line in _REFERENCE_VALUES})
__all__ = [name for name, _value, _line in _REFERENCE_VALUES]
