"""ctypes binding of libeagle_b200.so (include/eagle_b200.h).  No fallback: if the CUDA extension
has not been built, importing this module raises -- the product path never runs on the CPU."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# EAGLE_B200_LIBRARY: another build of the same library (the measurement build under tools/_variants/, a debug build)
LIB_PATH = os.environ.get("EAGLE_B200_LIBRARY") or os.path.join(_PKG, "libeagle_b200.so")

ABI_VERSION = 1
FIT_OK, FIT_FEW_POINTS, FIT_NO_MODEL, FIT_SKIPPED = 0, 1, 2, 3
KP_PY_INT, KP_NUMPY_INT, KP_FLOAT = 0, 1, 2
FIT_FIXED_K, FIT_CV2_COMPAT = 0, 1
NUM_LANDMARKS, ORDER_STRIDE, MODEL_H, MODEL_W = 57, 64, 540, 960


class NativeError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise NativeError(
        f"{LIB_PATH} is missing: build the CUDA extension first (python -m eagle_b200.build, or "
        "__graft_entry__.build()).  There is deliberately no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _d, _sz, _u64 = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_uint64

lib.egl_version.restype = _i
lib.egl_last_error.restype = C.c_char_p
lib.egl_sm_count.restype = _i
lib.egl_build_flags.restype = _i
lib.egl_preprocess_u8.argtypes = [_vp, _i, _i, _i, _sz, _sz, _vp, _vp]
lib.egl_preprocess_u8_letterbox.argtypes = [_vp, _i, _i, _i, _sz, _sz, _vp, _vp, _i, _i, _vp]
lib.egl_preprocess_u8_letterbox.restype = _i
lib.egl_upload_frames.argtypes = [_vp, _i, _sz, _vp, _i]
lib.egl_decode_heatmaps.argtypes = [_vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp]
lib.egl_decode_logits.argtypes = [_vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp]
lib.egl_synthesize_keypoints.argtypes = [_vp, _vp, _vp, _i, _i, _vp]
lib.egl_fit_homography.argtypes = [_vp, _vp, _vp, _i, _i, _i, _vp, _u64, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp]
lib.egl_select_homography.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp]
lib.egl_select_homography_chunk.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
lib.egl_project_points.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
lib.egl_pyramid_bytes.argtypes = [_i, _i, _i]
lib.egl_pyramid_bytes.restype = C.c_int64
lib.egl_gray_pyramid.argtypes = [_vp, _i, _i, _i, _sz, _sz, _i, _vp, _vp]
lib.egl_gray_pyramid_strided.argtypes = [_vp, _i, _i, _i, _sz, _sz, _i, _vp, _sz, _vp]
lib.egl_track_keypoints.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp]
lib.egl_filter_flow.argtypes = [_vp, _i, _i, _sz, _sz, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]
lib.egl_merge_keypoints.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]
lib.egl_calibrate_keypoints.argtypes = [_vp, _i, _i, _sz, _sz, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]
lib.egl_fit_homography_masked.argtypes = [_vp, _vp, _vp, _i, _i, _i, _vp, _u64, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.egl_commit_fit.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]
lib.egl_refine_keypoints.argtypes = [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]
lib.egl_fit_homography_subpixel.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _u64, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp]
for _name in ("egl_refine_keypoints", "egl_fit_homography_subpixel", "egl_gray_pyramid", "egl_track_keypoints", "egl_filter_flow", "egl_merge_keypoints", "egl_calibrate_keypoints",
              "egl_fit_homography_masked", "egl_commit_fit", "egl_refine_keypoints", "egl_fit_homography_subpixel"):
    getattr(lib, _name).restype = _i
for _name in ("egl_preprocess_u8", "egl_decode_heatmaps", "egl_decode_logits", "egl_synthesize_keypoints", "egl_fit_homography",
              "egl_select_homography", "egl_select_homography_chunk", "egl_project_points"):
    getattr(lib, _name).restype = _i

EXPORTS = ("egl_version", "egl_last_error", "egl_sm_count", "egl_build_flags", "egl_preprocess_u8", "egl_decode_heatmaps", "egl_decode_logits",
           "egl_synthesize_keypoints", "egl_fit_homography", "egl_select_homography", "egl_project_points", "egl_pyramid_bytes",
           "egl_gray_pyramid", "egl_track_keypoints", "egl_filter_flow", "egl_merge_keypoints", "egl_calibrate_keypoints",
           "egl_fit_homography_masked", "egl_commit_fit", "egl_refine_keypoints", "egl_fit_homography_subpixel",
           "egl_select_homography_chunk", "egl_upload_frames", "egl_gray_pyramid_strided", "egl_preprocess_u8_letterbox")

if lib.egl_version() != ABI_VERSION:
    raise NativeError(f"libeagle_b200.so has ABI {lib.egl_version()}, this package expects {ABI_VERSION}; rebuild")


def check(rc: int, what: str) -> None:
    """Raise on a non-zero return code of a C-ABI call (argument error > 0, CUDA error < 0)."""
    if rc != 0:
        raise NativeError(f"{what} failed (rc={rc}): {lib.egl_last_error().decode(errors='replace')}")
