"""ctypes access to tests/native/libhostcheck.so: the kernels' scalar core compiled for the host.

Test infrastructure: lets the CPU suite check the arithmetic the GPU will execute (same source,
g++ -ffp-contract=off) against the oracle.  Built on demand; never imported by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "host_check.cpp")
LIB = os.path.join(HERE, "native", "libhostcheck.so")
CORE = os.path.join(os.path.dirname(HERE), "eagle_b200", "csrc", "geometry_core.cuh")
FLOW_CORE = os.path.join(os.path.dirname(HERE), "eagle_b200", "csrc", "flow_core.cuh")
CASCADE_CORE = [os.path.join(os.path.dirname(HERE), "eagle_b200", "csrc", n) for n in ("cascade_core.cuh", "rho_hfunc.inc", "rho_lmstep.inc")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(LIB)) or any(os.path.getmtime(p) > os.path.getmtime(LIB) for p in (SRC, CORE, FLOW_CORE, *CASCADE_CORE))
        if stale:
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", LIB, SRC])
        _lib = C.CDLL(LIB)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def split(img_pts, world_pts):
    s = np.ascontiguousarray(img_pts, np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(world_pts, np.float32).reshape(-1, 2)
    return [np.ascontiguousarray(v) for v in (s[:, 0], s[:, 1], d[:, 0], d[:, 1])]


def fit_cv2(img_pts, world_pts, thr=5.0, confidence=0.995, max_iters=2000):
    sx, sy, dx, dy = split(img_pts, world_pts)
    H = np.zeros(9); mask = C.c_uint64(0); info = np.zeros(4, np.int32)
    st = lib().hc_fit_cv2(_p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float), len(sx),
                          C.c_double(thr), C.c_double(confidence), max_iters, _p(H, C.c_double), C.byref(mask), _p(info, C.c_int32))
    m = np.array([(mask.value >> i) & 1 for i in range(len(sx))], np.uint8)
    return st, H.reshape(3, 3), m, info


def fit_rho(img_pts, world_pts):
    """cascade_core.cuh rho_fit -> (inlier count (0 = no model), H float32 3x3, mask)."""
    sx, sy, dx, dy = split(img_pts, world_pts)
    H = np.zeros(9, np.float32); mask = C.c_uint64(0)
    n = lib().hc_fit_rho(_p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float), len(sx), _p(H, C.c_float),
                         C.byref(mask))
    m = np.array([(mask.value >> i) & 1 for i in range(len(sx))], np.uint8)
    return n, H.reshape(3, 3), m


def fit_lmeds(img_pts, world_pts, confidence=0.995):
    """cascade_core.cuh lmeds_fit -> (final inlier count (0 = no model), H 3x3, mask)."""
    sx, sy, dx, dy = split(img_pts, world_pts)
    H = np.zeros(9); mask = C.c_uint64(0); band = C.c_uint64(0)
    n = lib().hc_fit_lmeds(_p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float), len(sx),
                           C.c_double(confidence), _p(H, C.c_double), C.byref(mask), C.byref(band))
    m = np.array([(mask.value >> i) & 1 for i in range(len(sx))], np.uint8)
    fit_lmeds.band = np.array([(band.value >> i) & 1 for i in range(len(sx))], np.uint8)
    return n, H.reshape(3, 3), m


def fixedk_stage(img_pts, world_pts, K, hyp=None, seed=0, frame=0, thr=5.0):
    sx, sy, dx, dy = split(img_pts, world_pts)
    H = np.zeros(9, np.float64); mask = C.c_uint64(0); info = np.zeros(4, np.int32)
    hp = None
    if hyp is not None:
        hyp = np.ascontiguousarray(hyp, np.uint8); hp = _p(hyp, C.c_uint8)
    st = lib().hc_fixedk_stage(_p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float), len(sx), K, hp,
                               C.c_uint64(seed), C.c_uint64(frame), C.c_double(thr), _p(H, C.c_double), C.byref(mask), _p(info, C.c_int32))
    m = np.array([(mask.value >> i) & 1 for i in range(len(sx))], np.uint8)
    return st, H.reshape(3, 3), m, info


def refit(H, img_pts, world_pts, ransac_mask, thr=5.0):
    sx, sy, dx, dy = split(img_pts, world_pts)
    Hc = np.ascontiguousarray(H, np.float64).reshape(-1).copy(); fm = C.c_uint64(0)
    bits = sum(int(b) << i for i, b in enumerate(np.asarray(ransac_mask).ravel()))
    n = lib().hc_refit(_p(Hc, C.c_double), _p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float), len(sx),
                       C.c_uint64(bits), C.c_double(thr), C.byref(fm))
    return Hc.reshape(3, 3), np.array([(fm.value >> i) & 1 for i in range(len(sx))], np.uint8), n


def dlt4_f64(img4, world4):
    sx, sy, dx, dy = split(img4, world4)
    H = np.zeros(9)
    ok = lib().hc_dlt4_f64(_p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float), _p(H, C.c_double))
    return bool(ok), H.reshape(3, 3)


def check_subset(img4, world4):
    sx, sy, dx, dy = split(img4, world4)
    return bool(lib().hc_check_subset(_p(sx, C.c_float), _p(sy, C.c_float), _p(dx, C.c_float), _p(dy, C.c_float)))


def seeded_subset(seed, frame, K, h, N):
    idx = np.zeros(4, np.int32)
    lib().hc_seeded_subset(C.c_uint64(seed), C.c_uint64(frame), C.c_uint64(K), C.c_uint64(h), N, _p(idx, C.c_int))
    return idx


def postprocess(flat, score, hm_h, hm_w, img_w, img_h, conf=0.3):
    flat = np.ascontiguousarray(flat, np.int32); score = np.ascontiguousarray(score, np.float32)
    xy = np.zeros((57, 2), np.int32); order = np.zeros(64, np.uint8)
    n = lib().hc_postprocess(_p(flat, C.c_int32), _p(score, C.c_float), hm_h, hm_w, img_w, img_h, C.c_double(conf),
                             _p(xy, C.c_int32), _p(order, C.c_uint8))
    return xy, order[:n].copy()


def synthesize(xy, order, max_new=30):
    xy = np.ascontiguousarray(xy, np.int32).copy(); o = np.zeros(64, np.uint8); o[:len(order)] = order
    n = lib().hc_synthesize(_p(xy, C.c_int32), _p(o, C.c_uint8), len(order), max_new)
    return xy, o[:n].copy()


def project(H, pts):
    H = np.ascontiguousarray(H, np.float64).reshape(-1); pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    of = np.zeros_like(pts); oi = np.zeros(pts.shape, np.int64)
    lib().hc_project(_p(H, C.c_double), _p(pts, C.c_float), len(pts), _p(of, C.c_float), _p(oi, C.c_int64))
    return of, oi


def gray_hue(bgr):
    """(gray, hue) uint8 of an (..., 3) uint8 BGR array through the kernels' colour formulas."""
    a = np.ascontiguousarray(bgr, np.uint8).reshape(-1, 3)
    g = np.zeros(len(a), np.uint8); h = np.zeros(len(a), np.uint8)
    lib().hc_gray_hue(_p(a, C.c_uint8), len(a), _p(g, C.c_uint8), _p(h, C.c_uint8))
    return g.reshape(bgr.shape[:-1]), h.reshape(bgr.shape[:-1])


def build_pyramid(gray, max_level=2):
    """Pyramid levels of a gray image as the kernels lay them out (list of 2-D arrays)."""
    g = np.ascontiguousarray(gray, np.uint8)
    H, W = g.shape
    lib().hc_pyramid_bytes.restype = C.c_longlong
    buf = np.zeros(lib().hc_pyramid_bytes(H, W, max_level), np.uint8)
    n = lib().hc_build_pyramid(_p(g, C.c_uint8), H, W, max_level, _p(buf, C.c_uint8))
    out, off, h, w = [], 0, H, W
    for l in range(n):
        if l > 0:
            h, w = (h + 1) // 2, (w + 1) // 2
        out.append(buf[off:off + h * w].reshape(h, w).copy())
        off += (h * w + 15) // 16 * 16
    return out


def track(prev_gray, next_gray, pts, max_level=2, max_count=10, eps=0.03):
    """cv2.calcOpticalFlowPyrLK with winSize (15, 15) through the kernels' scalar tracker: (next_pts, status)."""
    a = np.ascontiguousarray(prev_gray, np.uint8); b = np.ascontiguousarray(next_gray, np.uint8)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(p); st = np.zeros(len(p), np.uint8)
    lib().hc_track(_p(a, C.c_uint8), _p(b, C.c_uint8), a.shape[0], a.shape[1], max_level, _p(p, C.c_float), len(p), max_count,
                   C.c_double(eps), _p(out, C.c_float), _p(st, C.c_uint8))
    return out, st


def pairwise_sum(v):
    a = np.ascontiguousarray(v, np.float32)
    lib().hc_pairwise_sum.restype = C.c_float
    return np.float32(lib().hc_pairwise_sum(_p(a, C.c_float), len(a)))
