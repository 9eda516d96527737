"""Oracle: Lucas-Kanade keypoint propagation between network frames (TEST INFRASTRUCTURE ONLY).

Restates ``CoordinateModel.calculate_optical_flow`` (eagle/models/coordinate_model.py:419-478) and the
OpenCV routines underneath it, integer step by integer step, so that the CUDA kernels have something
exact to be compared with:

  * ``cv2.cvtColor(BGR2GRAY)``              coordinate_model.py:281   -> gray_restated
  * ``cv2.calcOpticalFlowPyrLK``            coordinate_model.py:435   -> lk_track_restated
      (pyramid = pyrDown + REFLECT_101 border, Scharr derivatives, fixed-point window sampling with
       14-bit weights, float accumulation in the lane order of OpenCV's 128-bit SIMD build)
  * ``cv2.cvtColor(BGR2HSV)`` hue          coordinate_model.py:461,470 -> hue_restated
  * the z-score / hue filters               coordinate_model.py:441-476 -> filter_flow

OpenCV is a third-party dependency of the reference (opencv-python 4.11.0.86 in uv.lock; this image
has 4.13.0, SSE3 baseline build).  Every function here is pinned against the live library by
tests/test_oracle_optflow.py.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------------
# colour conversions
# --------------------------------------------------------------------------------------------------
def gray_restated(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(BGR2GRAY) for uint8: 15-bit fixed point, round half up."""
    b = bgr[..., 0].astype(np.int64); g = bgr[..., 1].astype(np.int64); r = bgr[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


_HDIV180 = np.zeros(256, np.int64)
_HDIV180[1:] = np.rint((180 << 12) / (6.0 * np.arange(1, 256))).astype(np.int64)


def hue_restated(bgr: np.ndarray) -> np.ndarray:
    """H channel of cv2.cvtColor(BGR2HSV) for uint8 (range 0..179), 12-bit fixed point."""
    b = bgr[..., 0].astype(np.int64); g = bgr[..., 1].astype(np.int64); r = bgr[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    vr = v == r
    vg = v == g
    h = np.where(vr, g - b, np.where(vg, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV180[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return h.astype(np.uint8)


def value_restated(bgr: np.ndarray) -> np.ndarray:
    """V channel of BGR2HSV = max(B, G, R)."""
    return bgr.max(axis=-1)


# --------------------------------------------------------------------------------------------------
# pyramid and derivatives
# --------------------------------------------------------------------------------------------------
def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def pyr_down_restated(g: np.ndarray) -> np.ndarray:
    """cv2.pyrDown for uint8: separable [1 4 6 4 1], exact integer sum, (s + 128) >> 8, REFLECT_101."""
    h, w = g.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int64)
    gi = g.astype(np.int64)
    xs = _reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)  # (ow, 5)
    rows = (gi[:, xs] * k).sum(-1)                                                 # (h, ow)
    ys = _reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)   # (oh, 5)
    out = (rows[ys, :] * k[None, :, None]).sum(1)
    return ((out + 128) >> 8).astype(np.uint8)


def scharr_restated(g: np.ndarray):
    """calcScharrDeriv (lkpyramid.cpp): (Ix, Iy) int16 with the 3/10/3 smoothing, REFLECT_101 inside the image."""
    h, w = g.shape
    gi = g.astype(np.int32)
    yu = _reflect101(np.arange(h) - 1, h); yd = _reflect101(np.arange(h) + 1, h)
    t0 = (gi[yu] + gi[yd]) * 3 + gi * 10
    t1 = gi[yd] - gi[yu]
    xl = _reflect101(np.arange(w) - 1, w); xr = _reflect101(np.arange(w) + 1, w)
    ix = t0[:, xr] - t0[:, xl]
    iy = (t1[:, xr] + t1[:, xl]) * 3 + t1 * 10
    return ix.astype(np.int16), iy.astype(np.int16)


def build_pyramid(g: np.ndarray, win: int, max_level: int):
    """buildOpticalFlowPyramid: levels until one is not larger than the window."""
    levels = [g]
    for _ in range(max_level):
        nxt = pyr_down_restated(levels[-1])
        if nxt.shape[1] <= win or nxt.shape[0] <= win:
            break
        levels.append(nxt)
    return levels


# --------------------------------------------------------------------------------------------------
# the tracker (LKTrackerInvoker, lkpyramid.cpp)
# --------------------------------------------------------------------------------------------------
W_BITS = 14
_FLT_SCALE = F32(1.0 / (1 << 20))


def _cv_round(x) -> int:
    """cvRound of a float: round half to even (SSE cvtss2si)."""
    return int(np.rint(np.float64(x)))


def _weights(a: F32, b: F32):
    one = F32(1.0)
    s = F32(1 << W_BITS)
    iw00 = _cv_round((one - a) * (one - b) * s)
    iw01 = _cv_round(a * (one - b) * s)
    iw10 = _cv_round((one - a) * b * s)
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _bordered(level: np.ndarray, win: int) -> np.ndarray:
    return np.pad(level, win, mode="reflect")  # numpy 'reflect' == BORDER_REFLECT_101


def _sample(img: np.ndarray, x0: int, y0: int, win: int, w, shift: int) -> np.ndarray:
    """(win, win) int32 of CV_DESCALE(bilinear with integer weights, shift); img indexable at [y0..y0+win] x [x0..x0+win]."""
    p = img[y0:y0 + win + 1, x0:x0 + win + 1].astype(np.int64)
    v = p[:-1, :-1] * w[0] + p[:-1, 1:] * w[1] + p[1:, :-1] * w[2] + p[1:, 1:] * w[3]
    return ((v + (1 << (shift - 1))) >> shift).astype(np.int32)


def _lane_sum(prod: np.ndarray, reduce_order: str = "movehl") -> F32:
    """Float accumulation of a (win, win) int32 product grid the way the SSE build does it.

    Columns 0..8*(win//8)-1 go through four float lanes (lane = column mod 4), row after row; the
    remaining columns are added one by one into a scalar, row after row; the lanes are then reduced
    as (l0 + l2) + (l1 + l3) -- the movehl/shuffle reduction; probing cv2 4.13 with
    OPTFLOW_LK_GET_MIN_EIGENVALS matches this order in 240/240 cases and (l0+l1)+(l2+l3) in only 224 --
    and added to the scalar."""
    win = prod.shape[1]
    nv = (win // 8) * 8
    lanes = np.zeros(4, F32)
    tail = F32(0.0)
    pf = prod.astype(F32)  # exact or rounded like cvtdq2ps / (float)int
    for y in range(prod.shape[0]):
        for x in range(0, nv, 4):
            lanes = lanes + pf[y, x:x + 4]
        for x in range(nv, win):
            tail = F32(tail + pf[y, x])
    if reduce_order == "hadd":
        s = F32(F32(lanes[0] + lanes[1]) + F32(lanes[2] + lanes[3]))
    else:
        s = F32(F32(lanes[0] + lanes[2]) + F32(lanes[1] + lanes[3]))
    return F32(tail + s)


def _mismatch_sum(diff: np.ndarray, dI: np.ndarray):
    """(ib1, ib2): float accumulation of diff*Ix, diff*Iy in the SSE build's order (see lk_track_restated)."""
    win = diff.shape[1]
    nv = (win // 8) * 8
    d = diff.astype(np.int64)
    px = d * dI[..., 0]; py = d * dI[..., 1]
    qb0 = np.zeros(4, F32); qb1 = np.zeros(4, F32)
    t1 = F32(0.0); t2 = F32(0.0)
    for y in range(diff.shape[0]):
        for x in range(0, nv, 8):
            # int32 pair sums (pixel k with pixel k+4), converted to float, then accumulated
            a = np.array([px[y, x] + px[y, x + 4], py[y, x] + py[y, x + 4], px[y, x + 1] + px[y, x + 5], py[y, x + 1] + py[y, x + 5]])
            b = np.array([px[y, x + 2] + px[y, x + 6], py[y, x + 2] + py[y, x + 6], px[y, x + 3] + px[y, x + 7], py[y, x + 3] + py[y, x + 7]])
            qb0 = qb0 + a.astype(F32)
            qb1 = qb1 + b.astype(F32)
        for x in range(nv, win):
            t1 = F32(t1 + F32(px[y, x])); t2 = F32(t2 + F32(py[y, x]))
    q = qb0 + qb1  # (X0, Y0, X1, Y1)
    ib1 = F32(t1 + F32(F32(q[0] + q[2]) + F32(0.0)))
    ib2 = F32(t2 + F32(F32(q[1] + q[3]) + F32(0.0)))
    return ib1, ib2


def lk_track_restated(prev_gray: np.ndarray, next_gray: np.ndarray, pts: np.ndarray, win: int = 15, max_level: int = 2,
                      max_count: int = 10, eps: float = 0.03, min_eig_threshold: float = 1e-4, reduce_order: str = "movehl"):
    """cv2.calcOpticalFlowPyrLK(prev, next, pts, None, winSize=(win,win), maxLevel, criteria=(EPS|COUNT, max_count, eps)).

    Returns (next_pts (N,2) float32, status (N,) uint8).  ``err`` is not restated (the reference ignores it)."""
    pts = np.asarray(pts, F32).reshape(-1, 2)
    n = len(pts)
    pyr_i = build_pyramid(prev_gray, win, max_level)
    pyr_j = build_pyramid(next_gray, win, max_level)
    top = min(len(pyr_i), len(pyr_j)) - 1
    eps2 = min(max(float(eps), 0.0), 10.0) ** 2
    max_count = min(max(int(max_count), 0), 100)
    next_pts = np.zeros((n, 2), F32)
    status = np.ones(n, np.uint8)
    half = F32((win - 1) * 0.5)
    for level in range(top, -1, -1):
        I = _bordered(pyr_i[level], win)
        J = _bordered(pyr_j[level], win)
        ix, iy = scharr_restated(pyr_i[level])
        D = np.zeros((I.shape[0], I.shape[1], 2), np.int16)  # BORDER_CONSTANT 0 outside the image
        D[win:-win, win:-win, 0] = ix; D[win:-win, win:-win, 1] = iy
        rows, cols = pyr_i[level].shape
        scale = F32(1.0 / (1 << level))
        for k in range(n):
            prev = pts[k] * scale
            if level == top:
                nxt = prev.copy()
            else:
                nxt = next_pts[k] * F32(2.0)
            next_pts[k] = nxt
            prev = prev - half
            ipx = int(np.floor(prev[0])); ipy = int(np.floor(prev[1]))
            if ipx < -win or ipx >= cols or ipy < -win or ipy >= rows:
                if level == 0:
                    status[k] = 0
                continue
            a = F32(prev[0] - F32(ipx)); b = F32(prev[1] - F32(ipy))
            w = _weights(a, b)
            Iw = _sample(I, ipx + win, ipy + win, win, w, W_BITS - 5)
            dIx = _sample(D[..., 0], ipx + win, ipy + win, win, w, W_BITS)
            dIy = _sample(D[..., 1], ipx + win, ipy + win, win, w, W_BITS)
            A11 = F32(_lane_sum(dIx * dIx, reduce_order) * _FLT_SCALE)
            A12 = F32(_lane_sum(dIx * dIy, reduce_order) * _FLT_SCALE)
            A22 = F32(_lane_sum(dIy * dIy, reduce_order) * _FLT_SCALE)
            Dt = F32(F32(A11 * A22) - F32(A12 * A12))
            dd = F32(A11 - A22)
            min_eig = F32(F32(F32(A22 + A11) - np.sqrt(F32(F32(dd * dd) + F32(F32(F32(4.0) * A12) * A12)))) / F32(2 * win * win))
            if float(min_eig) < min_eig_threshold or Dt < np.finfo(F32).eps:
                if level == 0:
                    status[k] = 0
                continue
            Dt = F32(F32(1.0) / Dt)
            nxt = nxt - half
            prev_delta = np.zeros(2, F32)
            dI = np.stack([dIx, dIy], -1).astype(np.int64)
            for j in range(max_count):
                inx = int(np.floor(nxt[0])); iny = int(np.floor(nxt[1]))
                if inx < -win or inx >= cols or iny < -win or iny >= rows:
                    if level == 0:
                        status[k] = 0
                    break
                a = F32(nxt[0] - F32(inx)); b = F32(nxt[1] - F32(iny))
                w = _weights(a, b)
                diff = _sample(J, inx + win, iny + win, win, w, W_BITS - 5) - Iw
                ib1, ib2 = _mismatch_sum(diff, dI)
                b1 = F32(ib1 * _FLT_SCALE); b2 = F32(ib2 * _FLT_SCALE)
                delta = np.array([F32(F32(F32(A12 * b2) - F32(A22 * b1)) * Dt), F32(F32(F32(A12 * b1) - F32(A11 * b2)) * Dt)], F32)
                nxt = nxt + delta
                next_pts[k] = nxt + half
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps2:
                    break
                if j > 0 and abs(float(F32(delta[0] + prev_delta[0]))) < 0.01 and abs(float(F32(delta[1] + prev_delta[1]))) < 0.01:
                    next_pts[k] = next_pts[k] - delta * F32(0.5)
                    break
                prev_delta = delta
            if status[k] and level == 0:
                # the error pass (always run: err is always requested by the Python binding) re-checks the
                # final position and marks the point lost when its window left the image
                fin = next_pts[k] - half
                fx = int(np.floor(fin[0])); fy = int(np.floor(fin[1]))
                if fx < -win or fx >= cols or fy < -win or fy >= rows:
                    status[k] = 0
    return next_pts, status


# --------------------------------------------------------------------------------------------------
# numpy's float32 reductions, restated (what np.mean / np.std / np.linalg.norm do to <= 128 floats)
# --------------------------------------------------------------------------------------------------
def pairwise_sum_f32(a: np.ndarray) -> F32:
    """numpy's pairwise summation for n <= 128 contiguous float32: < 8 elements sequentially from 0,
    else eight interleaved accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the
    remainder added sequentially."""
    a = np.asarray(a, F32)
    n = len(a)
    assert n <= 128
    if n < 8:
        s = F32(0.0) if n == 0 else a[0]
        for v in a[1:]:
            s = F32(s + v)
        return F32(s)
    r = a[:8].copy()
    i = 8
    while i + 8 <= n:
        r = r + a[i:i + 8]
        i += 8
    s = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
    for v in a[i:]:
        s = F32(s + v)
    return s


def move_stats_restated(new_pts: np.ndarray, prev_pts: np.ndarray):
    """coordinate_model.py:442-444 in explicit float32 steps: (move_amounts, mean, std + 1e-6)."""
    d = (np.asarray(new_pts, F32) - np.asarray(prev_pts, F32))
    sq = d * d
    move = np.sqrt((sq[:, 0] + sq[:, 1]).astype(F32)).astype(F32)
    n = F32(len(move))
    mean = F32(pairwise_sum_f32(move) / n)
    dev = (move - F32(pairwise_sum_f32(move) / n)).astype(F32)
    var = F32(pairwise_sum_f32(dev * dev) / n)
    std = F32(np.sqrt(var) + F32(1e-6))  # NEP 50: float32 scalar + python float stays float32
    return move, mean, std


def filter_flow(frame: np.ndarray, labels: list, prev_pts: np.ndarray, new_pts: np.ndarray, status: np.ndarray) -> dict:
    """coordinate_model.py:438-478: drop lost points, the z-score > 2 movers and the hue jumpers.

    ``labels`` are the keys of prev_keypoints in order.  Note :446 -- the reference indexes the label
    list with the position in the *status-filtered* arrays, so after a lost point every later point is
    emitted under an earlier label; that is reproduced here."""
    keep = np.asarray(status).reshape(-1) == 1
    new_f = np.asarray(new_pts, F32)[keep]
    prev_f = np.asarray(prev_pts, F32)[keep]
    out = {}
    if len(new_f) == 0:
        return out  # np.mean of an empty array is nan with a warning; the loop body never runs
    move, mean, std = move_stats_restated(new_f, prev_f)
    h, w = frame.shape[:2]

    def mean_hue(pt):
        x, y = (int(v) for v in pt.astype(int))
        x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1)
        grid = frame[max(0, y - 1):min(h, y + 2), max(0, x - 1):min(w, x + 2)]
        hue = hue_restated(grid)
        return float(hue.astype(np.int64).sum()) / hue.size

    for j in range(len(new_f)):
        z = F32(F32(move[j] - mean) / std)
        if z > 2:
            continue
        if abs(mean_hue(new_f[j]) - mean_hue(prev_f[j])) > 25:
            continue
        out[labels[j]] = tuple(new_f[j].astype(int))
    return out


def calculate_optical_flow_restated(frame, prev_gray, prev_keypoints: dict, curr_gray, win=15, max_level=2, max_count=10, eps=0.03) -> dict:
    """coordinate_model.py:419-478 end to end on restated pieces."""
    if prev_gray is None or curr_gray is None or prev_keypoints is None or len(prev_keypoints) == 0:
        return {}
    prev_pts = np.array(list(prev_keypoints.values()), dtype=F32)
    if prev_pts.ndim != 2 or prev_pts.shape[0] == 0 or prev_pts.shape[1] != 2:
        return {}
    new_pts, status = lk_track_restated(prev_gray, curr_gray, prev_pts, win, max_level, max_count, eps)
    return filter_flow(frame, list(prev_keypoints.keys()), prev_pts, new_pts, status)


# --------------------------------------------------------------------------------------------------
# brightness calibration (coordinate_model.py:520-555)
# --------------------------------------------------------------------------------------------------
def calibrate_keypoints_restated(frame: np.ndarray, keypoints: dict) -> dict:
    """Snap a dim keypoint (V < 150) to the brightest pixel of the 6x6 block frame[y-3:y+3, x-3:x+3].

    Quirks kept: the block is 6x6, not 7x7 (:543-544 use exclusive upper bounds); the offset of the
    arg-max is taken relative to OFFSET even when the block was clipped at the left/top edge (:551-552);
    ``grid_hsv[OFFSET, OFFSET]`` (:548) raises IndexError when the clipped block has fewer than 4 rows or
    columns, i.e. for a dim keypoint with x == 0 or y == 0 -- the exception propagates out of
    get_coordinates; adjusted values are numpy int64 (np.clip), untouched ones keep their type."""
    OFFSET, THR = 3, 150
    h, w = frame.shape[:2]
    out = {}
    for key, (x, y) in keypoints.items():
        if not (0 <= x < w and 0 <= y < h):
            out[key] = (x, y)
            continue
        if int(value_restated(frame[y, x])) >= THR:
            out[key] = (x, y)
            continue
        grid = frame[max(0, y - OFFSET):min(h, y + OFFSET), max(0, x - OFFSET):min(w, x + OFFSET)]
        v = value_restated(grid)
        if v.shape[0] <= OFFSET or v.shape[1] <= OFFSET:
            raise IndexError(f"index {OFFSET} is out of bounds (calibration block {v.shape} at the frame edge)")
        by, bx = np.unravel_index(np.argmax(v), v.shape)
        out[key] = (np.clip(x + bx - OFFSET, 0, w - 1), np.clip(y + by - OFFSET, 0, h - 1))
    return out


def calculate_optical_flow_cv2(frame, prev_gray, prev_keypoints: dict, curr_gray, win=15, max_level=2, max_count=10, eps=0.03) -> dict:
    """coordinate_model.py:419-478 with the library calls the reference itself makes (cv2.calcOpticalFlowPyrLK,
    cv2.cvtColor, numpy statistics): the CPU baseline of the propagation step, and a cross-check of the
    restated version above."""
    import cv2
    if prev_gray is None or curr_gray is None or prev_keypoints is None or len(prev_keypoints) == 0:
        return {}
    prev_points = np.array(list(prev_keypoints.values()), dtype=F32)
    if prev_points.ndim != 2 or prev_points.shape[0] == 0 or prev_points.shape[1] != 2:
        return {}
    new_points, status, _ = cv2.calcOpticalFlowPyrLK(prev_gray, curr_gray, prev_points, None, winSize=(win, win), maxLevel=max_level,
                                                     criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, max_count, eps))
    new_points = new_points[status[:, 0] == 1]
    prev_points = prev_points[status[:, 0] == 1]
    out = {}
    if len(new_points) == 0:
        return out
    move = np.linalg.norm(new_points - prev_points, axis=1)
    mean = np.mean(move)
    std = np.std(move) + 1e-6
    keys = list(prev_keypoints.keys())
    h, w = frame.shape[:2]

    def mean_hue(pt):
        x, y = pt.astype(int)
        x = np.clip(x, 0, w - 1); y = np.clip(y, 0, h - 1)
        grid = frame[max(0, y - 1):min(h, y + 2), max(0, x - 1):min(w, x + 2)]
        return np.mean(cv2.cvtColor(grid, cv2.COLOR_BGR2HSV)[:, :, 0])

    for j, (point, new_point) in enumerate(zip(prev_points, new_points)):
        if (move[j] - mean) / std > 2:
            continue
        if abs(mean_hue(new_point) - mean_hue(point)) > 25:
            continue
        out[keys[j]] = tuple(new_point.astype(int))
    return out
