"""Oracle: image->pitch homography fit (test infrastructure).

Two layers:

1. ``gather_correspondences`` / ``find_homography_cascade`` restate the reference's own statements
   (eagle/models/coordinate_model.py:335-357): build the on-plane point lists and call
   ``cv2.findHomography`` with the RANSAC -> RHO -> LMEDS cascade.  The arithmetic of that call
   lives in a third-party dependency that is NOT under /root/reference: opencv-python 4.11.0.86
   (uv.lock:992-993); this image carries cv2 4.13.0.

2. ``find_homography_restated`` restates OpenCV's published algorithm for
   ``findHomography(src, dst, RANSAC, thr)`` (calib3d: RANSACPointSetRegistrator::run,
   HomographyEstimatorCallback::{checkSubset,runKernel,computeError}, RANSACUpdateNumIters,
   HomographyRefineCallback + LMSolver) so that the same algorithm can also be driven with an
   EXPLICIT hypothesis table -- "both implementations fed the same seeded hypothesis set".  It is
   pinned against the live cv2 in tests/test_oracle_homography.py (H and mask, many seeds).
"""
from __future__ import annotations

import math

import cv2
import numpy as np

from .landmarks import NAME_TO_INDEX, NOT_ON_PLANE, WORLD

FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)


# --------------------------------------------------------------------------------------------
# layer 1: the reference's statements
# --------------------------------------------------------------------------------------------
def gather_correspondences(keypoints: dict):
    """coordinate_model.py:335-349: on-plane landmarks -> (img_pts, world_pts, used_labels)."""
    img_pts, world_pts, used_labels = [], [], []
    for label, (xi, yi) in keypoints.items():
        idx = NAME_TO_INDEX.get(label, -1)
        if idx in NOT_ON_PLANE:
            continue
        wx, wy, wz = WORLD[label]
        if wz != 0.0:
            continue
        img_pts.append([xi, yi])
        world_pts.append([wx, wy])
        used_labels.append(label)
    return (np.array(img_pts, dtype=np.float32).reshape(-1, 2),
            np.array(world_pts, dtype=np.float32).reshape(-1, 2), used_labels)


def find_homography_cascade(img_pts: np.ndarray, world_pts: np.ndarray):
    """coordinate_model.py:354-357.  Returns (H or None, mask or None, method used)."""
    H = mask = None
    method = None
    for method in [cv2.RANSAC, cv2.RHO, cv2.LMEDS]:
        H, mask = cv2.findHomography(img_pts, world_pts, method, 5.0 if method is cv2.RANSAC else None)
        if H is not None:
            break
    return H, mask, method


# --------------------------------------------------------------------------------------------
# layer 2: OpenCV internals, restated
# --------------------------------------------------------------------------------------------
class CvRNG:
    """cv::RNG (multiply-with-carry), seeded with 2**64-1 by RANSACPointSetRegistrator::run."""

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else a + self.next() % (b - a)


def _have_collinear(p: np.ndarray) -> bool:
    """haveCollinearPoints(ms, 4): only the LAST point is tested against earlier pairs (double)."""
    p = p.astype(np.float64)
    i = len(p) - 1
    for j in range(i):
        dx1 = p[j, 0] - p[i, 0]
        dy1 = p[j, 1] - p[i, 1]
        for k in range(j):
            dx2 = p[k, 0] - p[i, 0]
            dy2 = p[k, 1] - p[i, 1]
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3_rows1(p: np.ndarray, t) -> float:
    """determinant of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] in double, cv::determinant(Matx33d) order."""
    a, b, c = (p[t[0]].astype(np.float64), p[t[1]].astype(np.float64), p[t[2]].astype(np.float64))
    return (a[0] * (b[1] * 1.0 - 1.0 * c[1]) - a[1] * (b[0] * 1.0 - 1.0 * c[0]) + 1.0 * (b[0] * c[1] - b[1] * c[0]))


_TRIPLES = ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3))


def check_subset(src4: np.ndarray, dst4: np.ndarray) -> bool:
    """HomographyEstimatorCallback::checkSubset for a 4-point sample."""
    if _have_collinear(src4) or _have_collinear(dst4):
        return False
    negative = 0
    for t in _TRIPLES:
        negative += (_det3_rows1(src4, t) * _det3_rows1(dst4, t)) < 0
    return negative in (0, 4)


def run_kernel(src: np.ndarray, dst: np.ndarray):
    """HomographyEstimatorCallback::runKernel: normalised DLT in double.  src/dst (n,2) float32.

    Returns H (3,3) float64 with H[2,2] == 1, or None when a scale degenerates.
    """
    M = src.astype(np.float64)
    m = dst.astype(np.float64)
    n = len(M)
    cM = np.array([np.sum(M[:, 0]), np.sum(M[:, 1])]) / n
    cm = np.array([np.sum(m[:, 0]), np.sum(m[:, 1])]) / n
    # sequential sums as in the C loop (np.sum pairwise differs only beyond 1e-16; use math.fsum-free loop)
    sM = np.zeros(2)
    sm = np.zeros(2)
    cMx = cMy = cmx = cmy = 0.0
    for i in range(n):
        cmx += m[i, 0]; cmy += m[i, 1]; cMx += M[i, 0]; cMy += M[i, 1]
    cm = np.array([cmx / n, cmy / n]); cM = np.array([cMx / n, cMy / n])
    for i in range(n):
        sm[0] += abs(m[i, 0] - cm[0]); sm[1] += abs(m[i, 1] - cm[1])
        sM[0] += abs(M[i, 0] - cM[0]); sM[1] += abs(M[i, 1] - cM[1])
    if min(abs(sm[0]), abs(sm[1]), abs(sM[0]), abs(sM[1])) < DBL_EPSILON:
        return None
    sm = n / sm
    sM = n / sM
    invHnorm = np.array([[1.0 / sm[0], 0, cm[0]], [0, 1.0 / sm[1], cm[1]], [0, 0, 1]])
    Hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    LtL = np.zeros((9, 9))
    for i in range(n):
        x = (m[i, 0] - cm[0]) * sm[0]; y = (m[i, 1] - cm[1]) * sm[1]
        X = (M[i, 0] - cM[0]) * sM[0]; Y = (M[i, 1] - cM[1]) * sM[1]
        Lx = np.array([X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x])
        Ly = np.array([0, 0, 0, X, Y, 1, -y * X, -y * Y, -y])
        LtL += np.outer(Lx, Lx) + np.outer(Ly, Ly)
    LtL = np.triu(LtL) + np.triu(LtL, 1).T  # completeSymm (upper -> lower)
    _, _, V = cv2.eigen(LtL)  # rows = eigenvectors, eigenvalues descending (Jacobi, as cv::eigen)
    H0 = V[8].reshape(3, 3)
    H = (invHnorm @ H0) @ Hnorm2
    return H * (1.0 / H[2, 2])


def compute_error(src: np.ndarray, dst: np.ndarray, H: np.ndarray) -> np.ndarray:
    """HomographyEstimatorCallback::computeError: float32 arithmetic in OpenCV's operation order."""
    f = np.float32
    Hf = H.reshape(-1)[:8].astype(f)
    X = src[:, 0].astype(f); Y = src[:, 1].astype(f)
    x = dst[:, 0].astype(f); y = dst[:, 1].astype(f)
    with np.errstate(all="ignore"):
        ww = f(1.0) / ((Hf[6] * X + Hf[7] * Y) + f(1.0))
        dx = ((Hf[0] * X + Hf[1] * Y) + Hf[2]) * ww - x
        dy = ((Hf[3] * X + Hf[4] * Y) + Hf[5]) * ww - y
        return dx * dx + dy * dy


def find_inliers(src, dst, H, thresh: float):
    err = compute_error(src, dst, H)
    t = np.float32(thresh * thresh)
    with np.errstate(invalid="ignore"):
        mask = (err <= t)
    return int(mask.sum()), mask.astype(np.uint8)


def _cv_round(v: float) -> int:
    return int(np.rint(v))  # lrint: round half to even


def ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - math.pow(1.0 - ep, model_points)
    if denom < DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else _cv_round(num / denom)


def _refine_compute(h, src64, dst64, want_jac):
    """HomographyRefineCallback::compute -- the 9-parameter form (h33 is a free parameter).

    The cv2 binary of this image asserts ``J.cols == 9`` in fundam.cpp, i.e. the refinement runs
    over all nine entries and the result is rescaled by 1/h33 afterwards (older OpenCV releases
    optimised eight entries with h33 fixed to 1).
    """
    Mx = src64[:, 0]; My = src64[:, 1]
    ww = h[6] * Mx + h[7] * My + h[8]
    ww = np.where(np.abs(ww) > DBL_EPSILON, 1.0 / np.where(ww == 0, 1.0, ww), 0.0)
    xi = (h[0] * Mx + h[1] * My + h[2]) * ww
    yi = (h[3] * Mx + h[4] * My + h[5]) * ww
    err = np.empty(2 * len(Mx))
    err[0::2] = xi - dst64[:, 0]
    err[1::2] = yi - dst64[:, 1]
    if not want_jac:
        return err, None
    J = np.zeros((2 * len(Mx), 9))
    J[0::2, 0] = Mx * ww; J[0::2, 1] = My * ww; J[0::2, 2] = ww
    J[0::2, 6] = -Mx * ww * xi; J[0::2, 7] = -My * ww * xi; J[0::2, 8] = -ww * xi
    J[1::2, 3] = Mx * ww; J[1::2, 4] = My * ww; J[1::2, 5] = ww
    J[1::2, 6] = -Mx * ww * yi; J[1::2, 7] = -My * ww * yi; J[1::2, 8] = -ww * yi
    return err, J


def lm_refine(H: np.ndarray, src: np.ndarray, dst: np.ndarray, max_iters: int = 10, trace=None):
    """LMSolverImpl::run (calib3d levmarq.cpp) over HomographyRefineCallback; eps = FLT_EPSILON.

    Returns (H refined and rescaled so that H[2,2] == 1, iterations run).  ``trace`` (a list)
    receives one dict per iteration -- used to build golden vectors for the CUDA refit.
    """
    src64 = src.astype(np.float64); dst64 = dst.astype(np.float64)
    x = H.reshape(-1).astype(np.float64).copy()
    lx = 9
    epsx = epsf = FLT_EPSILON
    r, J = _refine_compute(x, src64, dst64, True)
    S = float(np.dot(r, r))
    A = J.T @ J
    v = J.T @ r
    D = np.diag(A).copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A.copy()
        Ap[np.diag_indices(lx)] += lam * D
        _, d = cv2.solve(Ap, v.reshape(-1, 1), flags=cv2.DECOMP_EIG)
        d = d.reshape(-1)
        xd = x - d
        rd, _ = _refine_compute(xd, src64, dst64, False)
        Sd = float(np.dot(rd, rd))
        temp_d = 2.0 * v - A @ d
        dS = float(np.dot(d, temp_d))
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if trace is not None:
            trace.append(dict(it=it, lam=lam, S=S, Sd=Sd, R=R, dmax=float(np.max(np.abs(d)))))
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < Rlo:
            t = float(np.dot(d, v))
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2.0
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                _, Ainv = cv2.invert(A, flags=cv2.DECOMP_EIG)
                maxval = max(DBL_EPSILON, float(np.max(np.abs(np.diag(Ainv)))))
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = _refine_compute(x, src64, dst64, True)
            A = J.T @ J
            v = J.T @ r
        it += 1
        proceed = it < max_iters and float(np.max(np.abs(d))) >= epsx and float(np.max(np.abs(r))) >= epsf
        if not proceed:
            break
    Hn = x.reshape(3, 3)
    scale = 1.0 / Hn[2, 2] if abs(Hn[2, 2]) > FLT_EPSILON else 1.0  # fundam.cpp scaleFor()
    return Hn * scale, it


def draw_subset(rng: CvRNG, count: int):
    """One pass of getSubset's index drawing: 4 distinct indices by rejection."""
    idx = []
    for _ in range(4):
        idx_i = rng.uniform(0, count)
        while idx_i in idx:
            idx_i = rng.uniform(0, count)
        idx.append(idx_i)
    return idx


def find_homography_restated(img_pts, world_pts, thresh: float = 5.0, hyp_table=None,
                             max_iters: int = 2000, confidence: float = 0.995,
                             adaptive: bool = True, minimal_solver=None, refit: str = "cv2",
                             recompute_mask: bool = True, return_info: bool = False):
    """findHomography(src, dst, RANSAC, thresh) restated.

    hyp_table: optional (K,4) int array of sample indices.  When given, hypotheses are taken from
        the table in order instead of from cv::RNG (rows failing checkSubset are skipped the way a
        rejected draw would be, but without consuming further table rows); with ``adaptive=False``
        every row is evaluated (fixed-K mode).
    minimal_solver: callable (src4, dst4) -> H or None; default = run_kernel (cv2's).
    refit: "cv2" = run_kernel on inliers + LM (what findHomography does).
    recompute_mask: return the inliers of the REFINED H (observed behaviour of cv2 4.13; see
        tests/test_oracle_homography.py) rather than the RANSAC best mask.
    """
    src = np.ascontiguousarray(img_pts, dtype=np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(world_pts, dtype=np.float32).reshape(-1, 2)
    count = len(src)
    solver = minimal_solver or run_kernel
    info = {"iters": 0, "best_hyp": -1, "ransac_mask": None, "n_models": 0}
    if thresh is None or thresh <= 0:
        thresh = 3.0

    def done(H, mask):
        return (H, mask, info) if return_info else (H, mask)

    if count < 4:
        return done(None, np.zeros((max(count, 0), 1), np.uint8))
    if count == 4:
        H = run_kernel(src, dst)
        if H is None:
            return done(None, np.zeros((4, 1), np.uint8))
        return done(H, np.ones((4, 1), np.uint8))

    rng = CvRNG()
    niters = max(max_iters, 1) if hyp_table is None else len(hyp_table)
    if hyp_table is not None and adaptive:
        niters = min(max(max_iters, 1), len(hyp_table))
    best_H, best_mask, max_good = None, None, 0
    it = 0
    while it < niters:
        if hyp_table is None:
            found = False
            for _ in range(10000):
                idx = draw_subset(rng, count)
                if check_subset(src[idx], dst[idx]):
                    found = True
                    break
            if not found:
                if it == 0:
                    return done(None, np.zeros((count, 1), np.uint8))
                break
        else:
            idx = [int(v) for v in hyp_table[it]]
            if not check_subset(src[idx], dst[idx]):
                it += 1
                continue
        H = solver(src[idx], dst[idx])
        if H is None:
            it += 1
            continue
        info["n_models"] += 1
        good, mask = find_inliers(src, dst, H, thresh)
        if good > max(max_good, 3):
            best_H, best_mask, max_good = H, mask, good
            info["best_hyp"] = it
            if adaptive:
                niters = ransac_update_num_iters(confidence, (count - good) / count, 4, niters)
        it += 1
    info["iters"] = it
    if max_good <= 0:
        return done(None, np.zeros((count, 1), np.uint8))
    info["ransac_mask"] = best_mask.copy()

    sel = best_mask.astype(bool)
    H = best_H
    if refit == "cv2":
        s_in, d_in = src[sel], dst[sel]
        Hk = run_kernel(s_in, d_in)
        if Hk is not None:
            H = Hk
        H, info["lm_iters"] = lm_refine(H, s_in, d_in, 10)
    mask = best_mask
    if recompute_mask:
        _, mask = find_inliers(src, dst, H, thresh)
    return done(H, mask.reshape(-1, 1))
