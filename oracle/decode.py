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
