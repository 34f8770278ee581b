"""ctypes binding of libsfmmatch.so (include/sfm_match.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFMM_LIB_PATH") or os.path.join(HERE, "csrc", "libsfmmatch.so")  # env override: kernel-variant experiments only

SFMM_OK, SFMM_EINVAL, SFMM_ENOMEM, SFMM_ECUDA, SFMM_ESTATE, SFMM_ERANGE, SFMM_ENODEVICE = 0, -1, -2, -3, -4, -5, -6
NORM_HAMMING, NORM_L2 = 0, 1
U8, F32 = 0, 1
FLOAT_AUTO, FLOAT_EXACT, FLOAT_TENSOR = 0, 1, 2
BINARY_AUTO, BINARY_POPC, BINARY_TENSOR = 0, 1, 2

#: numpy view of SfmDMatch == cv::DMatch
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])

#: every symbol include/sfm_match.h declares
EXPORTS = (
    "sfmm_version", "sfmm_default_config", "sfmm_create", "sfmm_destroy", "sfmm_last_error",
    "sfmm_set_descriptors", "sfmm_descriptor_blob", "sfmm_row_pitch", "sfmm_match_all_pairs",
    "sfmm_match_pairs", "sfmm_get_pair", "sfmm_match_pair", "sfmm_knn_pair", "sfmm_result_table",
    "sfmm_match_pairs_device", "sfmm_clear_results", "sfmm_get_stats",
    "sfmm_set_points", "sfmm_get_pair_points", "sfmm_save_table", "sfmm_load_table", "sfmm_image_rows",
    "sfmm_group_create", "sfmm_group_destroy", "sfmm_group_last_error", "sfmm_group_size", "sfmm_group_context",
    "sfmm_group_set_descriptors", "sfmm_group_match_all_pairs", "sfmm_group_match_pairs", "sfmm_group_get_pair",
    "sfmm_group_transfer_stats", "sfmm_share_table", "sfmm_shared_table_info",
)
#: every symbol include/sfm_features.h declares
FEATURE_EXPORTS = ("sfmm_orb_create", "sfmm_orb_destroy", "sfmm_orb_last_error", "sfmm_orb_detect_and_compute", "sfmm_orb_stats")


class SfmmConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("norm", C.c_int32), ("ratio", C.c_float),
                ("cross_check", C.c_int32), ("float_mode", C.c_int32), ("pair_batch", C.c_int32),
                ("binary_engine", C.c_int32)]


class SfmmStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("pairs_matched", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("last_match_ms", C.c_double), ("last_knn_ms", C.c_double),
                ("last_knn_work", C.c_double), ("last_knn_launches", C.c_int64), ("float_path", C.c_int64),
                ("tensor_kind", C.c_int64)]


class SfmmError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"libsfmmatch error {code}: {text}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sfm_danpipeline_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the matching path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    P = C.POINTER
    L.sfmm_version.restype = C.c_char_p
    L.sfmm_version.argtypes = []
    L.sfmm_default_config.restype = None
    L.sfmm_default_config.argtypes = [P(SfmmConfig)]
    L.sfmm_create.argtypes = [P(SfmmConfig), P(vp)]
    L.sfmm_destroy.restype = None
    L.sfmm_destroy.argtypes = [vp]
    L.sfmm_last_error.restype = C.c_char_p
    L.sfmm_last_error.argtypes = [vp]
    L.sfmm_set_descriptors.argtypes = [vp, i32, P(vp), P(i32), i32, P(sz), i32]
    L.sfmm_descriptor_blob.argtypes = [vp, P(vp), P(sz)]
    L.sfmm_row_pitch.restype = sz
    L.sfmm_row_pitch.argtypes = [i32, i32]
    L.sfmm_match_all_pairs.argtypes = [vp]
    L.sfmm_match_pairs.argtypes = [vp, vp, i64]
    L.sfmm_get_pair.argtypes = [vp, i32, i32, P(vp), P(i32)]
    L.sfmm_match_pair.argtypes = [vp, i32, i32, vp, i32, P(i32)]
    L.sfmm_knn_pair.argtypes = [vp, i32, i32, vp, vp]
    L.sfmm_result_table.argtypes = [vp, P(i64), P(vp), P(vp), P(vp), P(vp), P(i64)]
    L.sfmm_match_pairs_device.argtypes = [vp, vp, i64, vp, vp, i64, P(i64)]
    L.sfmm_clear_results.argtypes = [vp]
    L.sfmm_set_points.argtypes = [vp, i32, P(vp)]
    L.sfmm_get_pair_points.argtypes = [vp, i32, i32, P(vp), P(vp), P(i32)]
    L.sfmm_save_table.argtypes = [vp, C.c_char_p]
    L.sfmm_load_table.argtypes = [vp, C.c_char_p]
    L.sfmm_get_stats.argtypes = [vp, P(SfmmStats)]
    L.sfmm_image_rows.argtypes = [vp, i32, P(i32)]
    L.sfmm_orb_create.argtypes = [i32, P(vp)]
    L.sfmm_orb_destroy.restype = None
    L.sfmm_orb_destroy.argtypes = [vp]
    L.sfmm_orb_last_error.restype = C.c_char_p
    L.sfmm_orb_last_error.argtypes = [vp]
    L.sfmm_orb_detect_and_compute.argtypes = [vp, vp, i32, i32, sz, i32, vp, vp, i32, P(i32)]
    L.sfmm_orb_stats.argtypes = [vp, P(i64), P(C.c_double)]
    L.sfmm_share_table.argtypes = [vp, C.c_char_p]
    L.sfmm_shared_table_info.argtypes = [vp, C.c_char_p, sz, P(i64)]
    L.sfmm_group_create.argtypes = [P(SfmmConfig), i32, P(i32), P(vp)]
    L.sfmm_group_destroy.restype = None
    L.sfmm_group_destroy.argtypes = [vp]
    L.sfmm_group_last_error.restype = C.c_char_p
    L.sfmm_group_last_error.argtypes = [vp]
    L.sfmm_group_size.restype = i32
    L.sfmm_group_size.argtypes = [vp]
    L.sfmm_group_context.restype = vp
    L.sfmm_group_context.argtypes = [vp, i32]
    L.sfmm_group_set_descriptors.argtypes = [vp, i32, P(vp), P(i32), i32, P(sz), i32]
    L.sfmm_group_match_all_pairs.argtypes = [vp]
    L.sfmm_group_match_pairs.argtypes = [vp, vp, i64]
    L.sfmm_group_get_pair.argtypes = [vp, i32, i32, P(vp), P(i32)]
    L.sfmm_group_transfer_stats.argtypes = [vp, P(i64), P(i64)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:  # default
            fn.restype = C.c_int
    _lib = L
    return L
