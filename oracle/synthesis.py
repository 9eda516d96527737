"""Oracle: line-intersection keypoint synthesis (test infrastructure).

Restates eagle/models/coordinate_model.py:76-94 (_build_pitch_groups), :96-115 (_fit_line),
:117-138 (_intersect_lines) and :140-186 (_synthesize_keypoints_with_line_intersections).

One deliberate difference: the reference keeps each line family's members in a Python ``set`` of
label strings, so the order in which a family's points reach ``cv2.fitLine`` changes with
PYTHONHASHSEED (only the floating-point summation order inside fitLine depends on it).  Here the
members are kept in GROUND_TRUTH_POINTS dict order, one of the orders the reference can produce.
"""
from __future__ import annotations

import ctypes
import ctypes.util

import cv2
import numpy as np

from .landmarks import INDEX_TO_NAME, NAME_TO_INDEX, NOT_ON_PLANE, WORLD, WORLD_DICT_ORDER


def build_pitch_groups():
    """coordinate_model.py:76-94.  Returns (coord_to_label, x_groups, y_groups), insertion-ordered."""
    coord_to_label, x_groups, y_groups = {}, {}, {}
    for ch in WORLD_DICT_ORDER:
        label = INDEX_TO_NAME[ch]
        x, y, z = WORLD[label]
        if z != 0.0:
            continue
        xr = round(float(x), 2)
        yr = round(float(y), 2)
        if (xr, yr) not in coord_to_label:
            coord_to_label[(xr, yr)] = label
        x_groups.setdefault(xr, []).append(label)
        y_groups.setdefault(yr, []).append(label)
    return coord_to_label, x_groups, y_groups


_GROUPS = build_pitch_groups()

_LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_LIBM.cosf.restype = _LIBM.sinf.restype = ctypes.c_float
_LIBM.cosf.argtypes = _LIBM.sinf.argtypes = [ctypes.c_float]


def fit_line(points: np.ndarray):
    """coordinate_model.py:96-115."""
    if points is None or len(points) < 2:
        return None
    pts = points.astype(np.float32).reshape(-1, 1, 2)
    try:
        vx, vy, x0, y0 = cv2.fitLine(pts, cv2.DIST_L2, 0, 0.01, 0.01)
        vx = float(vx[0]); vy = float(vy[0]); x0 = float(x0[0]); y0 = float(y0[0])
        if abs(vx) + abs(vy) < 1e-6:
            return None
        return vx, vy, x0, y0
    except Exception:
        return None


def fit_line_restated(points: np.ndarray):
    """cv::fitLine(DIST_L2) for 2-D float points (imgproc linefit.cpp, fitLine2D_wods), restated.

    Moments accumulate in double (the squares/products are formed in float first), the angle is
    ``t = (float)atan2(2*dxy, dx2 - dy2) / 2`` and the direction is ``(cos(t), sin(t))`` on the
    FLOAT t, i.e. the C library's cosf/sinf (a double cos() rounded to float differs by 1 ulp in
    about 1 % of cases).
    Pinned against the live cv2.fitLine in tests/test_oracle_synthesis.py.
    """
    p = points.astype(np.float32).reshape(-1, 2)
    n = len(p)
    f = np.float32
    x = y = x2 = y2 = xy = 0.0
    for i in range(n):
        px, py = p[i, 0], p[i, 1]
        x += float(px); y += float(py)
        x2 += float(f(px * px)); y2 += float(f(py * py)); xy += float(f(px * py))
    w = float(f(n))
    x /= w; y /= w; x2 /= w; y2 /= w; xy /= w
    dx2 = x2 - x * x
    dy2 = y2 - y * y
    dxy = xy - x * y
    t = f(f(np.arctan2(2 * dxy, dx2 - dy2)) / f(2))
    return float(_LIBM.cosf(float(t))), float(_LIBM.sinf(float(t))), float(f(x)), float(f(y))


def intersect_lines(line1, line2):
    """coordinate_model.py:117-138."""
    if line1 is None or line2 is None:
        return None
    vx1, vy1, x01, y01 = line1
    vx2, vy2, x02, y02 = line2
    det = vx1 * (-vy2) - vy1 * (-vx2)
    if abs(det) < 1e-8:
        return None
    rhs = np.array([x02 - x01, y02 - y01], dtype=np.float64)
    A = np.array([[vx1, -vx2], [vy1, -vy2]], dtype=np.float64)
    try:
        t, _ = np.linalg.solve(A, rhs)
        return float(x01 + t * vx1), float(y01 + t * vy1)
    except Exception:
        return None


def synthesize(keypoints: dict, min_points_per_line: int = 2, max_new_points: int = 30, line_fn=fit_line) -> dict:
    """coordinate_model.py:140-186.  ``keypoints`` maps label -> (xi, yi); returns the merged dict."""
    coord_to_label, x_groups, y_groups = _GROUPS
    detected = {k: v for k, v in keypoints.items() if NAME_TO_INDEX.get(k, -1) not in NOT_ON_PLANE}
    lines_y = {}
    for y_val, labels in y_groups.items():
        pts = [detected[lbl] for lbl in labels if lbl in detected]
        if len(pts) >= min_points_per_line:
            line = line_fn(np.array(pts, dtype=np.float32))
            if line is not None:
                lines_y[y_val] = line
    lines_x = {}
    for x_val, labels in x_groups.items():
        pts = [detected[lbl] for lbl in labels if lbl in detected]
        if len(pts) >= min_points_per_line:
            line = line_fn(np.array(pts, dtype=np.float32))
            if line is not None:
                lines_x[x_val] = line
    added = {}
    for y_val, ly in lines_y.items():
        for x_val, lx in lines_x.items():
            label = coord_to_label.get((round(float(x_val), 2), round(float(y_val), 2)))
            if not label or label in keypoints:
                continue
            pt = intersect_lines(ly, lx)
            if pt is None:
                continue
            added[label] = (int(round(pt[0])), int(round(pt[1])))
            if len(added) >= max_new_points:
                break
        if len(added) >= max_new_points:
            break
    if added:
        return {**keypoints, **added}
    return keypoints
