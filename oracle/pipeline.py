"""Oracle: the per-frame geometry path end to end (test infrastructure; also the CPU baseline).

Restates the frame loop of ``CoordinateModel.get_coordinates``
(eagle/models/coordinate_model.py:277-415) for the cadence the accelerated path covers -- every frame
is a keypoint frame (``num_keypoint_detection == fps``, keypoint interval 1 at :206) and the homography
is refreshed every ``homography_interval`` frames or after a failed attempt (:205,:333,:350-367); the
benchmark uses interval 1 -- with
  * the network forward replaced by "heatmaps are given" (decode = keypoint_hrnet.py:583-594),
  * ``detect_objects`` replaced by "boxes are given",
  * the optical-flow rescue for frames with < 4 model keypoints (:285-311) left out: such frames
    simply get no fit and reuse the previous homography, which is what the reference does whenever
    the rescue finds nothing to add.
Cross-frame state kept exactly as the reference keeps it: the last successful homography
(:363-367, :375-378) and ``prev_keypoints`` (what is emitted as "Keypoints", :359-362,:415).
"""
from __future__ import annotations

import numpy as np

from . import decode as _decode
from . import homography as _hom
from . import project as _proj
from . import synthesis as _syn


def frame_time(i: int, fps: int) -> str:
    """coordinate_model.py:415."""
    return f"{i // fps // 60:02d}:{i // fps % 60:02d}"


def get_coordinates(heatmaps, objects_per_frame, width: int, height: int, fps: int = 1,
                    keypoint_conf: float = 0.3, synthesis: bool = True, trace: list | None = None,
                    fit=None, num_homography: int | None = None) -> dict:
    """Run the restated path over F frames.

    heatmaps: (F, 57, h, w) float32 (any array-like indexable by frame).
    objects_per_frame: list of detect_objects()-shaped dicts.
    fit: optional callable (img_pts, world_pts) -> (H, mask); default is the reference's
         cv2.findHomography cascade (coordinate_model.py:354-357).
    trace: optional list that receives one dict per frame with the intermediate values the
         reference does not return (decoded/synthesised keypoints, point lists, H, mask, raw
         projections) -- this is what the parity tests compare the CUDA path against.
    """
    res = {}
    homography_matrix = None
    prev_homography_matrix = None
    prev_keypoints = {}
    compute_homography = False
    # :205 -- homography cadence; num_homography=None means "every frame" (interval 1)
    homography_interval = 1 if num_homography is None else max(1, int(fps / max(1, num_homography)))
    F = len(objects_per_frame)
    for i in range(F):
        t = {} if trace is not None else None
        keypoints = _decode.decode_frame(np.asarray(heatmaps[i]), width, height, keypoint_conf)
        if t is not None:
            t["decoded"] = dict(keypoints)
        if synthesis and len(keypoints) >= 2:  # :326-327
            keypoints = _syn.synthesize(keypoints)
        if t is not None:
            t["synthesised"] = dict(keypoints)
        prev_keypoints = keypoints  # :330
        objects = objects_per_frame[i]
        attempt = (i % homography_interval == 0) or compute_homography  # :333
        img_pts, world_pts, used_labels = _hom.gather_correspondences(keypoints)
        if t is not None:
            t.update(img_pts=img_pts, world_pts=world_pts, used_labels=used_labels, H=None, mask=None, attempted=attempt)
        if attempt and len(img_pts) < 4:
            compute_homography = True  # :350-352
        if attempt and len(img_pts) >= 4:
            if fit is None:
                new_H, mask, _ = _hom.find_homography_cascade(img_pts, world_pts)
            else:
                new_H, mask = fit(img_pts, world_pts)
            if new_H is not None:
                if mask is not None and mask.size == len(used_labels):
                    keypoints = {k: v for k, v, m in zip(used_labels, img_pts.tolist(), mask.flatten()) if m}
                    prev_keypoints = keypoints
                homography_matrix = new_H
                prev_homography_matrix = homography_matrix
                compute_homography = False
                if t is not None:
                    t.update(H=new_H.copy(), mask=None if mask is None else mask.copy())
            else:
                compute_homography = True  # :366-367
        H_use = homography_matrix if homography_matrix is not None else prev_homography_matrix
        indiv, raw = _proj.project_objects(objects, H_use)
        bounds = _proj.boundaries(width, height, H_use)
        if t is not None:
            t.update(H_use=None if H_use is None else H_use.copy(), proj_raw=raw)
            trace.append(t)
        res[i] = {"Coordinates": indiv, "Time": frame_time(i, fps), "Keypoints": prev_keypoints,
                  "Boundaries": bounds}
    return res
