"""ctypes binding of libfacfake.so (include/facfake.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfacfake.so")

FF_OK, FF_ERR_BAD_ARG, FF_ERR_SHAPE, FF_ERR_CUDA, FF_ERR_STATE = 0, -1, -2, -3, -4
FF_X_NCHW_F32, FF_X_NHWC_U8 = 0, 2
FF_COMPUTE_BF16, FF_COMPUTE_FP32 = 0, 1
FF_REDUCE_REFERENCE, FF_REDUCE_SOFTMAX_MEAN, FF_REDUCE_REFERENCE_PROBS = 0, 1, 2

# every symbol include/facfake.h declares: (name, restype, argtypes)
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = [
    ("ff_cvit_create", _i, [C.POINTER(_vp), _i, _i, _i]),
    ("ff_resvitkan_create", _i, [C.POINTER(_vp), _i, _i, _i]),
    ("ff_cvit_ggca_create", _i, [C.POINTER(_vp), _i, _i, _i]),
    ("ff_cvit_destroy", None, [_vp]),
    ("ff_last_error", C.c_char_p, [_vp]),
    ("ff_cvit_load_weight", _i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    ("ff_cvit_finalize_weights", _i, [_vp]),
    ("ff_cvit_unused_keys", C.c_char_p, [_vp]),
    ("ff_preprocess_crops", _i, [_vp, C.POINTER(_vp), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _i, _vp, _vp, _vp]),
    ("ff_cvit_forward", _i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    ("ff_video_scores", _i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    ("ff_cvit_predict", _i, [_vp, _vp, _i, C.POINTER(C.c_int32), _vp, _i, _i, _vp, _vp, _vp]),
    ("ff_cvit_predict_host", _i, [_vp, _vp, C.POINTER(C.c_int32), _i, _i, _vp, _vp]),
    ("ff_cvit_launch_count", _i64, [_vp]),
    ("ff_cvit_debug_activation", _i64, [_vp, _vp, _i, _vp, _i, _i, _vp, _i64, _vp]),
    ("ff_cvit_set_tuning", _i, [_vp, _i]),
    ("ff_cvit_set_profiling", _i, [_vp, _i]),
    ("ff_cvit_get_profile", _i, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    ("ff_blazeface_create", _i, [C.POINTER(_vp), _i, _i]),
    ("ff_blazeface_destroy", None, [_vp]),
    ("ff_blazeface_last_error", C.c_char_p, [_vp]),
    ("ff_blazeface_load_weight", _i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    ("ff_blazeface_finalize", _i, [_vp]),
    ("ff_blazeface_predict", _i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    ("ff_blazeface_nms", _i, [_vp, _vp, _i, C.c_float, C.c_float, _vp, _vp, _vp]),
    ("ff_blazeface_nms_lists", _i, [_vp, _vp, _vp, _i, C.c_float, _vp, _vp, _vp]),
    ("ff_blazeface_tile_frames", _i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    ("ff_blazeface_frame_faces", _i, [_vp, _vp, _i, _i, _i, C.c_float, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    ("ff_blazeface_launch_count", _i64, [_vp]),
    ("ff_s3d_create", _i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    ("ff_s3d_destroy", None, [_vp]),
    ("ff_s3d_last_error", C.c_char_p, [_vp]),
    ("ff_s3d_load_weight", _i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    ("ff_s3d_finalize", _i, [_vp]),
    ("ff_s3d_forward", _i, [_vp, _vp, _i, _i, _vp, _vp]),
    ("ff_s3d_debug_activation", _i64, [_vp, _vp, _i, _i, _i, _vp, _i64, _vp]),
    ("ff_s3d_launch_count", _i64, [_vp]),
]

_lib = None


def load() -> C.CDLL:
    """Load libfacfake.so (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make` (nvcc, sm_100a). "
            "fac_fake_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
