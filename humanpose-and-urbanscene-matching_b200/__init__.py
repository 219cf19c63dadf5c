"""B200-native (sm_100a) matching hot path of gilbeckers/HumanPose-and-UrbanScene-Matching.

Layout
------
csrc/        hand-written CUDA kernels + the C ABI  -> libhpusm.so (see include/hpusm.h)
_lib.py      ctypes binding of the C ABI (fails loudly when the library or a GPU is missing)
engine.py    batched entry points over torch CUDA tensors (buffer carriers only)
retrieval.py query-vs-database drivers, sharded over ranks with torch.distributed
posedb.py / imagedb.py  packed pose / image-descriptor databases streamed from disk to HBM
dropin/      modules with the reference's own names and signatures (urbanscene.features, ...)

The directory name carries hyphens, so import it with
``importlib.import_module("humanpose-and-urbanscene-matching_b200")`` or through the ``hpusm`` alias
module at the repository root.
"""
from . import _lib  # noqa: F401
from . import synth  # noqa: F401
from . import engine, retrieval, posedb, imagedb, dropin  # noqa: F401
from .engine import (  # noqa: F401
    knn2_hamming,
    knn2_hamming_segmented,
    knn2_l2,
    ratio_filter,
    gather_matches,
    ransac_homography_batch,
    ransac_score_batch,
    perspective_transform,
    affine_lsq_batch,
    procrustes_batch,
    pose_match_batch,
    pose_match_pairs,
    pose_topk,
    launch_count,
)

__all__ = [
    "knn2_hamming", "knn2_hamming_segmented", "knn2_l2", "ratio_filter", "gather_matches",
    "ransac_homography_batch", "ransac_score_batch", "perspective_transform", "affine_lsq_batch",
    "procrustes_batch", "pose_match_batch", "pose_match_pairs", "pose_topk", "launch_count",
]
