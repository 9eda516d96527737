"""The C-ABI library loads and exports every function include/eagle_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "eagle_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(egl_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    fns = declared_functions()
    for name in ("egl_preprocess_u8", "egl_decode_heatmaps", "egl_synthesize_keypoints", "egl_fit_homography",
                 "egl_select_homography", "egl_project_points", "egl_version", "egl_last_error", "egl_sm_count"):
        assert name in fns


def test_library_exports_every_declared_symbol():
    from eagle_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/eagle_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared_functions()
    assert _native.lib.egl_version() == _native.ABI_VERSION


def test_argument_errors_do_not_need_a_gpu():
    """Null pointers / bad shapes are rejected before any CUDA call, with a message."""
    from eagle_b200 import _native as N
    rc = N.lib.egl_decode_heatmaps(None, 1, 135, 240, 1920, 1080, 0.3, None, None, None, None, None, None)
    assert rc == 1 and b"null pointer" in N.lib.egl_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = N.lib.egl_decode_heatmaps(p, 1, 3, 5, 1920, 1080, 0.3, p, p, p, p, p, None)  # 15 elements: not a multiple of 4
    assert rc == 2
    rc = N.lib.egl_fit_homography(p, p, p, 1, 7, 10, None, 0, 5.0, 0.995, p, p, p, p, p, None)
    assert rc == 4 and b"unknown mode" in N.lib.egl_last_error()


def test_product_never_imports_the_oracle():
    """No module under eagle_b200/ (nor bench.py's GPU arm imports) may depend on oracle/."""
    pkg = os.path.join(ROOT, "eagle_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_assembler_reproduces_reference_value_types():
    """Host-side dict assembly for the sparse cadence: kp_src says which Python type the reference holds each
    keypoint value in (int tuple / numpy-int tuple / float list) -- json.dumps(default=float) depends on it."""
    import json

    import numpy as np

    from eagle_b200 import _native as N
    from eagle_b200.coordinate_model import assemble_frames
    from eagle_b200.pitch import LANDMARK_NAMES
    xy = np.zeros((2, 57, 2), np.int32); order = np.zeros((2, 64), np.uint8); count = np.zeros((2, 2), np.int32); src = np.zeros((2, 64), np.uint8)
    order[0, :3] = (42, 13, 5); count[0] = (3, 3)
    xy[0, 42] = (10, 20); xy[0, 13] = (30, 40); xy[0, 5] = (50, 60)
    src[0, 42] = N.KP_PY_INT; src[0, 13] = N.KP_NUMPY_INT; src[0, 5] = N.KP_FLOAT
    objs = [{"Player": {}, "Goalkeeper": {}}, {"Player": {7: {"BBox": [1, 2, 3, 4], "Confidence": 0.5, "Bottom_center": [2, 4]}}, "Goalkeeper": {}}]
    res = assemble_frames(objs, 25, 0, xy, order, count, None, None, None, None, np.array([-1, -1], np.int32),
                          np.zeros((2, 1, 2), np.int64), np.zeros((2, 1), np.uint8), np.full((2, 4), np.nan), kp_src=src)
    kp = res[0]["Keypoints"]
    assert list(kp) == [LANDMARK_NAMES[42], LANDMARK_NAMES[13], LANDMARK_NAMES[5]]
    a, b, c = kp.values()
    assert a == (10, 20) and type(a) is tuple and type(a[0]) is int
    assert type(b) is tuple and isinstance(b[0], np.int64) and tuple(int(v) for v in b) == (30, 40)
    assert c == [50.0, 60.0] and type(c) is list and type(c[0]) is float
    assert json.dumps(kp, default=float) == '{"%s": [10, 20], "%s": [30.0, 40.0], "%s": [50.0, 60.0]}' % tuple(kp)
    assert res[1]["Keypoints"] == {} and res[1]["Boundaries"] == [None] * 4
    assert res[1]["Coordinates"]["Player"][7]["Transformed_Coordinates"] is None and res[1]["Time"] == "00:00"


def test_binding_argument_counts_match_the_header():
    """Every ctypes prototype in eagle_b200/_native.py has as many arguments as the declaration in the header."""
    from eagle_b200 import _native
    txt = open(os.path.join(ROOT, "include", "eagle_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for name, params in re.findall(r"\b(egl_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        argtypes = getattr(_native.lib, name).argtypes
        if argtypes is None:
            assert n == 0, f"{name}: header has {n} parameters, binding declares none"
        else:
            assert len(argtypes) == n, f"{name}: header has {n} parameters, binding {len(argtypes)}"
    # new propagation entry points reject null pointers before touching CUDA
    N = _native
    assert N.lib.egl_gray_pyramid(None, 1, 64, 64, 192, 64 * 192, 2, None, None) == 1
    assert N.lib.egl_track_keypoints(None, 64, 64, 2, None, None, None, 1, 0, 1, 1, 10, 0.03, None, None, None) == 1
    assert N.lib.egl_pyramid_bytes(1080, 1920, 2) == 2721600 and N.lib.egl_pyramid_bytes(16, 16, 2) == 256
    assert N.lib.egl_gray_pyramid(None, 0, 64, 64, 192, 64 * 192, 2, None, None) == 0   # empty batch: no-op
