"""Oracle: heatmap decode and keypoint post-processing (test infrastructure).

Restates, statement for statement,
  * ``KeypointModel.get_keypoints``            eagle/models/keypoint_hrnet.py:583-594
  * the keypoint post-processing block         eagle/models/coordinate_model.py:229-248
    (identical copies at :255-274 and :500-517)
with the network forward replaced by "the heatmaps are given".
"""
from __future__ import annotations

from collections import Counter

import numpy as np

from .landmarks import INDEX_TO_NAME


def get_keypoints(batch_heatmaps: np.ndarray):
    """keypoint_hrnet.py:583-594 on an (N, C, H, W) float32 array.

    Returns list (per frame) of lists of (channel, x_n, y_n, score): flat argmax (first maximum in
    row-major order), score = max value as Python float, x_n = x/(W-1), y_n = y/(H-1) in float64,
    kept when score > 0.01.
    """
    batch_coords = []
    for heatmaps in batch_heatmaps:
        coords = []
        for i in range(heatmaps.shape[0]):
            heatmap = np.asarray(heatmaps[i])
            H, W = heatmap.shape
            y, x = np.unravel_index(np.argmax(heatmap), heatmap.shape)
            score = float(heatmap[y, x])
            x_n = x / max(1, W - 1)
            y_n = y / max(1, H - 1)
            if score > 0.01:
                coords.append((i, x_n, y_n, score))
        batch_coords.append(coords)
    return batch_coords


def postprocess(kp_list, width: int, height: int, keypoint_conf: float = 0.3) -> dict:
    """coordinate_model.py:229-248: confidence filter, scale to image pixels, de-duplicate.

    ``kp_list`` is one frame's list from :func:`get_keypoints`.  Returns ``{label: (xi, yi)}`` in
    the reference's insertion order.
    """
    tmp = {}
    for i_lab, x_n, y_n, score in kp_list:
        if score < keypoint_conf:
            continue
        label = INDEX_TO_NAME[i_lab]
        xi = int(x_n * width)
        yi = int(y_n * height)
        tmp[label] = (xi, yi, score, i_lab)
    vals = list(tmp.values())
    coords = [v[:2] for v in vals]
    counts = Counter(coords)
    coords_to_label = {}
    for kk, vv in tmp.items():
        if counts[vv[:2]] == 1:
            coords_to_label[vv[:2]] = kk
        else:
            if vv[2] == max([xv[2] for xv in vals if xv[:2] == vv[:2]]):
                coords_to_label[vv[:2]] = kk
    return {coords_to_label[kp]: kp for kp in coords_to_label}


def decode_frame(heatmaps: np.ndarray, width: int, height: int, keypoint_conf: float = 0.3) -> dict:
    """get_keypoints + postprocess for a single (C, H, W) frame."""
    return postprocess(get_keypoints(heatmaps[None])[0], width, height, keypoint_conf)


def refine_subpixel(heatmap: np.ndarray, img_w: int, img_h: int) -> np.ndarray:
    """Specification of egl_refine_keypoints (an extension; the reference has no sub-pixel step): per channel a
    parabola through the arg-max and its two neighbours on each axis, float32 arithmetic in this order,
    offset clamped to +-0.5 and 0 on the border; returns (57, 2) float32 image positions."""
    F32 = np.float32
    C, h, w = heatmap.shape
    out = np.zeros((C, 2), F32)
    for c in range(C):
        m = heatmap[c].astype(F32)
        flat = int(np.argmax(m))
        y, x = divmod(flat, w)
        v = m[y, x]
        dx = dy = F32(0.0)
        if 0 < x < w - 1:
            l, r = m[y, x - 1], m[y, x + 1]
            den = F32(F32(v - l) + F32(v - r))
            if den > 0:
                dx = F32(min(max(F32(F32(F32(0.5) * F32(r - l)) / den), F32(-0.5)), F32(0.5)))
        if 0 < y < h - 1:
            u, d = m[y - 1, x], m[y + 1, x]
            den = F32(F32(v - u) + F32(v - d))
            if den > 0:
                dy = F32(min(max(F32(F32(F32(0.5) * F32(d - u)) / den), F32(-0.5)), F32(0.5)))
        out[c, 0] = F32(F32(F32(F32(x) + dx) / F32(w - 1)) * F32(img_w))
        out[c, 1] = F32(F32(F32(F32(y) + dy) / F32(h - 1)) * F32(img_h))
    return out
