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


def get_coordinates_propagated(frames, heatmaps, objects_per_frame, fps: int, num_homography: int = 1,
                               num_keypoint_detection: int = 1, calibration: bool = False, keypoint_conf: float = 0.3,
                               fit=None, trace: list | None = None, library_calls: bool = False) -> dict:
    """The full frame loop of ``get_coordinates`` (coordinate_model.py:205-417) including the sparse
    keypoint cadence: the network's heatmaps are decoded every ``keypoint_interval`` frames (:206,:216)
    and the landmarks are carried in between by Lucas-Kanade flow (:313-322), with the model fallback
    (:316-320), the < 4 keypoint rescues (:287-311) and optional brightness calibration (:328-329).

    frames: sequence of (H, W, 3) uint8 BGR; heatmaps: indexable by frame (only the frames the loop
    actually asks the network about are read).  Built from the restated pieces in oracle/optflow.py,
    each pinned against the live library; library_calls=True swaps in the cv2 calls themselves."""
    from . import optflow as _of

    height, width = frames[0].shape[:2]
    homography_interval = max(1, int(fps / max(1, num_homography)))
    keypoint_interval = max(1, int(fps / max(1, num_keypoint_detection)))
    detect = lambda j: _decode.decode_frame(np.asarray(heatmaps[j]), width, height, keypoint_conf)  # :480-518
    if library_calls:  # the reference's own cv2 calls instead of their restatements (CPU baseline; same results)
        import cv2
        gray = lambda j: cv2.cvtColor(frames[j], cv2.COLOR_BGR2GRAY)
        flow = lambda frame, pg, pk, cg: _of.calculate_optical_flow_cv2(frame, pg, pk, cg)
    else:
        gray = lambda j: _of.gray_restated(frames[j])
        flow = lambda frame, pg, pk, cg: _of.calculate_optical_flow_restated(frame, pg, pk, cg)
    prev_gray = None
    prev_keypoints = {}
    res = {}
    mem = {j: detect(j) for j in range(0, len(frames), keypoint_interval)}  # :216-274
    compute_homography = False
    homography_matrix = None
    prev_homography_matrix = None
    for i in range(len(frames)):
        frame = frames[i]
        t = {} if trace is not None else None
        curr_gray = gray(i)
        if i == 0 or i % keypoint_interval == 0:
            keypoints = mem.get(i)
            if len(keypoints) < 4:
                if i == 0:  # :288-307
                    j = None
                    for j in range(1, len(frames)):
                        next_gray = gray(j)
                        if j not in mem:
                            mem[j] = detect(j)
                        if len(mem[j]) >= 4:
                            prev_keypoints = mem[j]
                            break
                    if len(prev_keypoints) > 0:
                        for jj in range(j - 1, -1, -1):
                            pg = gray(jj)
                            flowed = flow(frames[jj], pg, prev_keypoints, next_gray)
                            prev_keypoints = flowed if len(flowed) > 0 else prev_keypoints
                            mem[jj] = {**prev_keypoints, **mem.get(jj, {})}
                            next_gray = pg
                else:  # :308-311
                    keypoints = {**keypoints, **flow(frame, prev_gray, prev_keypoints, curr_gray)}
        else:
            flowed = flow(frame, prev_gray, prev_keypoints, curr_gray)
            if t is not None:
                t["flowed"] = dict(flowed)
            if len(flowed) < 4:  # :316-320
                if i not in mem:
                    mem[i] = detect(i)
                keypoints = {**mem[i], **flowed}
            else:
                keypoints = {**flowed, **mem.get(i, {})}
        keypoints = {**keypoints, **mem.get(i, {})}  # :324
        if len(keypoints) >= 2:
            keypoints = _syn.synthesize(keypoints)
        if calibration:
            keypoints = _of.calibrate_keypoints_restated(frame, keypoints)
        if t is not None:
            t["synthesised"] = dict(keypoints)
        prev_keypoints = keypoints
        prev_gray = curr_gray
        objects = objects_per_frame[i]
        attempt = (i % homography_interval == 0) or compute_homography
        if t is not None:
            t.update(attempted=attempt, H=None, mask=None)
        if attempt:
            img_pts, world_pts, used_labels = _hom.gather_correspondences(keypoints)
            if len(img_pts) < 4:
                compute_homography = True
            else:
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
                    compute_homography = True
        H_use = homography_matrix if homography_matrix is not None else prev_homography_matrix
        indiv, raw = _proj.project_objects(objects, H_use)
        bounds = _proj.boundaries(width, height, H_use)
        if t is not None:
            t.update(H_use=None if H_use is None else H_use.copy(), proj_raw=raw)
            trace.append(t)
        res[i] = {"Coordinates": indiv, "Time": frame_time(i, fps), "Keypoints": prev_keypoints, "Boundaries": bounds}
    return res
