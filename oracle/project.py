"""Oracle: foot-point projection and visible-pitch boundaries (test infrastructure).

Restates eagle/models/coordinate_model.py:369-392 (per-object projection), :396-414 (boundaries)
and :32-44 (find_x_at_y).  ``cv2.perspectiveTransform`` is third-party (opencv-python 4.11.0.86 in
uv.lock:992-993, 4.13.0 in this image); ``perspective_transform_restated`` restates its arithmetic
and is pinned against the live call in tests/test_oracle_project.py.
"""
from __future__ import annotations

import cv2
import numpy as np

from .landmarks import PITCH_X_MAX, PITCH_Y_MAX

DBL_EPSILON = float(np.finfo(np.float64).eps)


def perspective_transform_restated(pts: np.ndarray, H: np.ndarray) -> np.ndarray:
    """cv::perspectiveTransform for float32 (n,2) points and a float64 3x3 matrix.

    Per point in double: w = x*h6 + y*h7 + h8; |w| > eps ? (x*h0 + y*h1 + h2)/w : 0, stored as
    float32 (OpenCV multiplies by the reciprocal: w = 1/w first).
    """
    h = np.asarray(H, dtype=np.float64).reshape(-1)
    p = np.asarray(pts, dtype=np.float32).reshape(-1, 2).astype(np.float64)
    x, y = p[:, 0], p[:, 1]
    w = x * h[6] + y * h[7] + h[8]
    ok = np.abs(w) > DBL_EPSILON
    winv = np.where(ok, 1.0 / np.where(ok, w, 1.0), 0.0)
    out = np.stack([(x * h[0] + y * h[1] + h[2]) * winv, (x * h[3] + y * h[4] + h[5]) * winv], axis=1)
    return out.astype(np.float32)


def find_x_at_y(pt1, pt2, y_target):
    """coordinate_model.py:32-44 (raises ZeroDivisionError exactly where the reference does)."""
    x1, y1 = pt1
    x2, y2 = pt2
    m = (y2 - y1) / (x2 - x1)
    c = y1 - m * x1
    return (y_target - c) / m


def project_objects(objects: dict, H_use):
    """coordinate_model.py:369-392 for one frame.  ``objects`` is a detect_objects()-shaped dict.

    Returns (indiv dict, list of pre-truncation float32 projections in iteration order).
    """
    indiv = {}
    raw = []
    for class_name, class_dict in objects.items():
        for obj_id, obj_dict in class_dict.items():
            bottom_center = obj_dict["Bottom_center"]
            bbox_coords = np.array(obj_dict["BBox"], dtype=np.uint16).tolist()
            conf = obj_dict["Confidence"]
            if H_use is None:
                curr = {int(obj_id): {"BBox": bbox_coords, "Confidence": conf, "Transformed_Coordinates": None,
                                      "Image_Bottom_center": bottom_center}}
            else:
                coords = np.array([[bottom_center]], dtype=np.float32)
                t_f = cv2.perspectiveTransform(coords, H_use)[0]
                raw.append(t_f[0].copy())
                t_i = t_f.astype(int)
                tx, ty = t_i[0, 0], t_i[0, 1]
                if tx < 0 or tx > PITCH_X_MAX or ty < 0 or ty > PITCH_Y_MAX:
                    curr = {int(obj_id): {"BBox": bbox_coords, "Confidence": conf, "Transformed_Coordinates": None,
                                          "Image_Bottom_center": bottom_center}}
                else:
                    curr = {int(obj_id): {"BBox": bbox_coords, "Confidence": conf,
                                          "Transformed_Coordinates": t_i.tolist()[0]}}
            if class_name not in indiv:
                indiv[class_name] = curr
            else:
                indiv[class_name].update(curr)
    return indiv, raw


def boundaries(width: int, height: int, H_use):
    """coordinate_model.py:396-414: [bottom_left, top_left, top_right, bottom_right] or 4x None."""
    top_left = top_right = bottom_left = bottom_right = None
    if H_use is not None:
        def proj(x, y):
            return cv2.perspectiveTransform(np.array([[[x, y]]], dtype=np.float32), H_use)[0].astype(int)[0].tolist()
        top_left = proj(0, 0)
        top_right = proj(width, 0)
        bottom_left = proj(0, height)
        bottom_right = proj(width, height)
    out = [None, None, None, None]
    if top_left is not None:
        try:
            top_left = (find_x_at_y(top_left, bottom_left, PITCH_Y_MAX), PITCH_Y_MAX)
            top_right = (find_x_at_y(top_right, bottom_right, PITCH_Y_MAX), PITCH_Y_MAX)
            bottom_left = (find_x_at_y(bottom_left, top_left, 0), 0)
            bottom_right = (find_x_at_y(bottom_right, top_right, 0), 0)
            out = [bottom_left, top_left, top_right, bottom_right]
        except Exception:
            pass
    return out
