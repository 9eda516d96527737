"""Oracle (test infrastructure): ctypes wrapper of oracle/ransac_f32.c, the independent C mirror of the
CUDA fixed-K hypothesis arithmetic, plus the cv2-faithful refit of its winner."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_c, homography

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
    return _lib


def seeded_table(seed: int, frame: int, K: int, n: int) -> np.ndarray:
    t = np.zeros((K, 4), np.uint8)
    lib().orc_seeded_table(C.c_uint64(seed), C.c_uint64(frame), K, n, t.ctypes.data_as(C.POINTER(C.c_uint8)))
    return t


def fixedk_frame(img_pts, world_pts, table, thr: float = 5.0):
    """Returns dict(best_index, count, mask (n,) uint8 in the normalised test, h (8,) float32, norm (6,))."""
    s = np.ascontiguousarray(img_pts, np.float32).reshape(-1, 2); d = np.ascontiguousarray(world_pts, np.float32).reshape(-1, 2)
    X, Y, x, y = (np.ascontiguousarray(v) for v in (s[:, 0], s[:, 1], d[:, 0], d[:, 1]))
    table = np.ascontiguousarray(table, np.uint8)
    bi = C.c_int(-1); bc = C.c_int(0); bm = C.c_uint64(0)
    h = np.zeros(8, np.float32); nm = np.zeros(6, np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rc = lib().orc_fixedk_frame(fp(X), fp(Y), fp(x), fp(y), len(X), table.ctypes.data_as(C.POINTER(C.c_uint8)), len(table),
                                C.c_float(np.float32(1.0 / thr)), C.byref(bi), C.byref(bc), C.byref(bm), fp(h), fp(nm))
    mask = np.array([(bm.value >> i) & 1 for i in range(len(X))], np.uint8)
    return dict(rc=rc, best_index=bi.value, count=bc.value, mask=mask, h=h, norm=nm)


def denormalise(h: np.ndarray, norm: np.ndarray, thr: float) -> np.ndarray:
    """Normalised hypothesis -> image->pitch homography (double), H[2,2] = 1."""
    cX, cY, sX, sY, cx, cy = (float(v) for v in norm)
    Hn = np.array([[h[0], h[1], h[2]], [h[3], h[4], h[5]], [h[6], h[7], 1.0]], np.float64)
    Ts = np.array([[sX, 0, -cX * sX], [0, sY, -cY * sY], [0, 0, 1.0]])
    Tdi = np.array([[thr, 0, cx], [0, thr, cy], [0, 0, 1.0]])
    H = Tdi @ Hn @ Ts
    return H / H[2, 2]


def fit_fixedk(img_pts, world_pts, table, thr: float = 5.0):
    """Whole fixed-K fit as the CUDA path does it: C-mirror hypothesis stage, then OpenCV's tail
    (inliers of the winner in cv2's float scoring -> runKernel -> LM -> mask from the refined H)."""
    r = fixedk_frame(img_pts, world_pts, table, thr)
    if r["best_index"] < 0:
        return None, None, r
    H0 = denormalise(r["h"], r["norm"], thr)
    src = np.ascontiguousarray(img_pts, np.float32).reshape(-1, 2); dst = np.ascontiguousarray(world_pts, np.float32).reshape(-1, 2)
    n0, m0 = homography.find_inliers(src, dst, H0, thr)
    if n0 < 4:
        return None, None, r
    sel = m0.astype(bool)
    Hk = homography.run_kernel(src[sel], dst[sel])
    H, _ = homography.lm_refine(Hk if Hk is not None else H0, src[sel], dst[sel], 10)
    _, mask = homography.find_inliers(src, dst, H, thr)
    return H, mask, r
