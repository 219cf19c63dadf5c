"""ctypes binding of libhpusm.so (the C ABI declared in include/hpusm.h).

There is deliberately no fallback: if the shared library is missing the import raises, and every
compute call raises ``HpusmError`` when the library reports a failure (e.g. no CUDA device).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhpusm.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = -1, -2, -3, -4
L2_AUTO, L2_SIMT, L2_TC_INT, L2_TC_SPLIT, L2_TC_I8, L2_TC_Q16 = 0, 1, 2, 3, 4, 5
POSE_DETAIL = 8 + 27 + 44


class HpusmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libhpusm error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libhpusm.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "humanpose-and-urbanscene-matching_b200/csrc/build.sh. There is no CPU fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int
_sz = ctypes.c_size_t
_f32 = ctypes.c_float
_f64 = ctypes.c_double
_u64 = ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/hpusm.h one to one
SIGNATURES = {
    "hpusm_version": (_i32, []),
    "hpusm_async_status": (_i32, []),
    "hpusm_last_error": (ctypes.c_char_p, []),
    "hpusm_device_count": (_i32, []),
    "hpusm_launch_count": (_i64, []),
    "hpusm_profile_enable": (_i32, [_i32]),
    "hpusm_profile_collect": (_i32, [_p, _p]),
    "hpusm_knn2_hamming_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "hpusm_knn2_hamming": (_i32, [_p, _i64, _p, _i64, _i32, _p, _p, _p, _sz, _p]),
    "hpusm_knn2_hamming_segmented_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "hpusm_knn2_hamming_segmented": (_i32, [_p, _i64, _p, _i64, _p, _i64, _i32, _f64, _i32, _p, _p, _p, _sz, _p]),
    "hpusm_knn2_l2_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "hpusm_knn2_l2": (_i32, [_p, _i64, _p, _i64, _i32, _i32, _p, _p, _p, _sz, _p]),
    "hpusm_knn2_l2_last_mode": (_i32, []),
    "hpusm_knn2_l2_last_fallback_rows": (_i64, []),
    "hpusm_knn2_merge_f32": (_i32, [_p, _p, _i32, _i64, _p, _p, _p]),
    "hpusm_knn2_merge_i32": (_i32, [_p, _p, _i32, _i64, _p, _p, _p]),
    "hpusm_ratio_filter": (_i32, [_p, _p, _i64, _i32, _f64, _p, _p, _p]),
    "hpusm_gather_matches_workspace_bytes": (_sz, [_i64]),
    "hpusm_gather_matches": (_i32, [_p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "hpusm_segment_matches": (_i32, [_p, _p, _i64, _i64, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "hpusm_validate_homography_batch": (_i32, [_p, _i64, _p, _p]),
    "hpusm_ransac_homography_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "hpusm_ransac_homography_batch": (_i32, [_p, _p, _p, _i64, _i64, _f32, _i32, _f64, _i32, _u64, _i64, _i32, _p, _p, _p,
                                             _p, _p, _sz, _p]),
    "hpusm_ransac_score_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "hpusm_ransac_score_batch": (_i32, [_p, _p, _p, _i64, _i64, _f32, _i32, _i32, _u64, _i64, _p, _p, _sz, _p]),
    "hpusm_perspective_transform": (_i32, [_p, _i64, _p, _p, _p]),
    "hpusm_affine_lsq_batch": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p]),
    "hpusm_procrustes_batch": (_i32, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "hpusm_pose_match_workspace_bytes": (_sz, [_i64]),
    "hpusm_pose_match_batch": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "hpusm_pose_match_batch_f64": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p]),
    "hpusm_pose_match_pairs": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p]),
    "hpusm_scene_score_batch": (_i32, [_p, _p, _p, _p, _p, _p, _i32, _p, _i64, _i32, _f64, _f64, _u64, _p, _p, _p, _p]),
    "hpusm_feature_scaling_batch": (_i32, [_p, _i64, _i32, _p, _p]),
    "hpusm_pose_topk_workspace_bytes": (_sz, [_i64, _i32]),
    "hpusm_pose_topk": (_i32, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "hpusm_pose_topk_merge": (_i32, [_p, _p, _i64, _i32, _p, _p, _p]),
    "hpusm_orb_workspace_bytes": (_sz, [_i32, _i32, _i32, _f64]),
    "hpusm_orb_compute": (_i32, [_p, _i32, _i32, _i64, _i32, _f64, _p, _p, _i64, _p, _p, _sz, _p]),
    "hpusm_orb_detect_workspace_bytes": (_sz, [_i32, _i32, _i32, _f64]),
    "hpusm_orb_detect_and_compute": (_i32, [_p, _i32, _i32, _i64, _i32, _f64, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _p, _p, _p,
                                            _p, _p, _p, _sz, _p]),
    "hpusm_microbench_popc": (_i32, [_i64, _i32, _i64, _p, _p]),
    "hpusm_microbench_lop3": (_i32, [_i64, _i32, _i64, _p, _p]),
    "hpusm_microbench_ffma": (_i32, [_i32, _i64, _i32, _i64, _p, _p]),
    "hpusm_microbench_minmax": (_i32, [_i32, _i64, _i32, _i64, _p, _p]),
    "hpusm_microbench_mma": (_i32, [_i32, _i32, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    msg = lib.hpusm_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code):
    if code != OK:
        raise HpusmError(code, last_error())
    return code
