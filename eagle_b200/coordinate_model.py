"""Drop-in host adapter: the reference's processor-level call signatures over the CUDA path.

Mirrors ``eagle.models.coordinate_model.CoordinateModel`` for the geometry path:

    model = CoordinateModel(keypoint_model=..., detect_objects=...)
    coords = model.get_coordinates(frames, fps, num_homography=1, num_keypoint_detection=1)

returns the dict ``Processor.__init__`` takes (eagle/processor.py:65-68) and docs/data.md:20-42
documents: ``{frame_idx: {"Coordinates", "Time", "Keypoints", "Boundaries"}}``.

What runs where
  * the two networks are NOT part of this package: ``keypoint_model`` is any callable mapping a
    (N,3,540,960) float32 CUDA tensor to (N,57,135,240) heatmaps (the reference's HRNet-W48 +
    sigmoid, keypoint_hrnet.py:565-573); ``detect_objects`` is any callable frame -> boxes dict
    (coordinate_model.py:557-628, YOLO + BoT-SORT);
  * preprocessing, heatmap decode, keypoint synthesis, homography fit, cadence selection and
    projection are CUDA kernels (eagle_b200/csrc) called through the C ABI;
  * this module only moves data and assembles Python dicts.

Cadence: the homography cadence (``num_homography``, retry after a failed fit, reuse of the
previous H) is reproduced exactly.  With a keypoint interval of 1 and every frame decoding >= 4
landmarks the frames are independent and go through GeometryPath in one batch; any other case
(sparse keypoint cadence, brightness calibration, a frame with < 4 landmarks that the reference
rescues by optical flow) goes through eagle_b200.propagation.PropagatedPath, which reproduces the
reference's Lucas-Kanade propagation (coordinate_model.py:277-330, 419-478, 520-555) on the GPU.
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np
import torch

from . import _native as N
from .engine import GeometryEngine
from .pitch import LANDMARK_NAMES, OFF_PLANE, PITCH_LENGTH_M, PITCH_WIDTH_M
from .boxes import max_objects, objects_to_arrays

BATCH = 4  # reference's HRNet batch (coordinate_model.py:20); only the chunking of the network calls


def frame_time(i: int, fps: int) -> str:
    """coordinate_model.py:415."""
    return f"{i // fps // 60:02d}:{i // fps % 60:02d}"


def _bbox_list(bbox):
    """``np.array(obj["BBox"], dtype=np.uint16).tolist()`` (coordinate_model.py:373) without the numpy round trip
    when the box already is a list of in-range Python ints (the detector clips person boxes to the frame)."""
    if type(bbox) is list and len(bbox) == 4:
        a, b, c, d = bbox
        if type(a) is int and type(b) is int and type(c) is int and type(d) is int and 0 <= a <= 65535 and 0 <= b <= 65535 \
                and 0 <= c <= 65535 and 0 <= d <= 65535:
            return [a, b, c, d]
    return np.array(bbox, dtype=np.uint16).tolist()


def assemble_frames_py(objects_per_frame, fps: int, first_index: int, kp_xy, kp_order, kp_count, used_mask, inlier_mask,
                       status, attempted, h_index, coords_i, in_bounds, bounds, kp_src=None) -> dict:
    """The readable statement of the dict assembly (coordinate_model.py:359-362, 369-392, 405-415) from the arrays
    the kernels produced (all numpy, already on the host).  ``assemble_frames`` -- the C extension built from
    csrc/assemble.c -- must return exactly this (tests/test_assemble.py); it is what the product calls.

    With ``kp_src`` (keypoint propagation) the keypoint sets are already what the reference holds in
    ``prev_keypoints`` at the end of each frame -- the inlier commit ran on the device -- and kp_src says
    which Python type each value has there (EGL_KP_*), so that json.dump(default=float) prints the same."""
    res = {}
    off = set(OFF_PLANE)
    names = LANDMARK_NAMES
    xy_l = np.asarray(kp_xy).tolist(); order_l = np.asarray(kp_order).tolist(); n_l = np.asarray(kp_count)[:, 0].tolist()
    hi_l = np.asarray(h_index).tolist(); ci_l = np.asarray(coords_i).tolist(); ib_l = np.asarray(in_bounds).tolist()
    bd_l = np.asarray(bounds, dtype=np.float64).tolist()
    if kp_src is None:
        inl_l = np.asarray(inlier_mask).tolist(); st_l = np.asarray(status).tolist(); att_l = np.asarray(attempted).tolist()
    else:
        src_l = np.asarray(kp_src).tolist()
        i64 = np.int64
    for k, objects in enumerate(objects_per_frame):
        i = first_index + k
        # --- "Keypoints": inliers as float lists when this frame's fit was used, else all keypoints
        chans = order_l[k][:n_l[k]]
        xy = xy_l[k]
        if kp_src is not None:
            src = src_l[k]
            keypoints = {}
            for c in chans:
                t = src[c]
                x, y = xy[c]
                keypoints[names[c]] = (x, y) if t == N.KP_PY_INT else ([float(x), float(y)] if t == N.KP_FLOAT else (i64(x), i64(y)))
        elif att_l[k] and st_l[k] == N.FIT_OK:
            inl = inl_l[k]
            keypoints = {names[c]: [float(xy[c][0]), float(xy[c][1])] for c in chans if c not in off and (inl >> c) & 1}
        else:
            keypoints = {names[c]: (xy[c][0], xy[c][1]) for c in chans}
        # --- "Coordinates"
        have_h = hi_l[k] >= 0
        indiv = {}
        p = 0
        ci = ci_l[k]; ib = ib_l[k]
        for class_name, class_dict in objects.items():
            for obj_id, obj in class_dict.items():
                if have_h and ib[p]:
                    curr = {int(obj_id): {"BBox": _bbox_list(obj["BBox"]), "Confidence": obj["Confidence"],
                                          "Transformed_Coordinates": ci[p]}}
                else:
                    curr = {int(obj_id): {"BBox": _bbox_list(obj["BBox"]), "Confidence": obj["Confidence"],
                                          "Transformed_Coordinates": None, "Image_Bottom_center": obj["Bottom_center"]}}
                p += 1
                if class_name not in indiv:
                    indiv[class_name] = curr
                else:
                    indiv[class_name].update(curr)
        # --- "Boundaries"
        b = bd_l[k]
        if have_h and b[0] == b[0]:
            boundaries = [(b[0], 0), (b[1], PITCH_WIDTH_M), (b[2], PITCH_WIDTH_M), (b[3], 0)]
        else:
            boundaries = [None, None, None, None]
        res[i] = {"Coordinates": indiv, "Time": frame_time(i, fps), "Keypoints": keypoints, "Boundaries": boundaries}
    return res


_OFF_PLANE_MASK = sum(1 << c for c in OFF_PLANE)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def assemble_frames(objects_per_frame, fps: int, first_index: int, kp_xy, kp_order, kp_count, used_mask, inlier_mask,
                    status, attempted, h_index, coords_i, in_bounds, bounds, kp_src=None, out: dict | None = None) -> dict:
    """Dict assembly through the C extension (csrc/assemble.c): a few microseconds per frame instead of 50-150.
    Same arguments and result as assemble_frames_py; ``out`` lets a caller append the frames of successive chunks
    to one dict.  The cyclic garbage collector is paused meanwhile: the records are millions of small containers,
    none of them cyclic, and generation-2 passes over them would cost more than the assembly itself."""
    import gc

    from . import _assemble
    res = {} if out is None else out
    objects_per_frame = objects_per_frame if type(objects_per_frame) is list else list(objects_per_frame)
    if kp_src is None:
        fitted = _c((np.asarray(attempted) != 0) & (np.asarray(status) == N.FIT_OK), np.uint8)
        inl, src = _c(inlier_mask, np.int64), None
    else:
        fitted, inl, src = None, None, _c(kp_src, np.uint8)
    was = gc.isenabled()
    gc.disable()
    try:
        _assemble.assemble(res, objects_per_frame, int(fps), int(first_index), LANDMARK_NAMES, _c(kp_xy, np.int32), _c(kp_order, np.uint8),
                           _c(kp_count, np.int32), inl, fitted, _c(h_index, np.int32), _c(coords_i, np.int64), _c(in_bounds, np.uint8),
                           _c(bounds, np.float64), src, _OFF_PLANE_MASK, int(PITCH_WIDTH_M), _bbox_list, np.int64, int(N.KP_PY_INT),
                           int(N.KP_FLOAT))
    finally:
        if was:
            gc.enable()
    return res


class GeometryPath:
    """decode -> synthesis -> fit -> cadence -> projection for frames whose heatmaps and boxes exist."""

    def __init__(self, device="cuda:0", keypoint_conf: float = 0.3, synthesis: bool = True, fit_mode: int = N.FIT_CV2_COMPAT,
                 max_iters: int = 2000, thr: float = 5.0):
        self.engine = GeometryEngine(device)
        self.keypoint_conf = keypoint_conf
        self.synthesis = synthesis
        self.fit_mode = fit_mode
        self.max_iters = max_iters
        self.thr = thr

    def run_device(self, heatmaps: torch.Tensor, foot: torch.Tensor, count: torch.Tensor, width: int, height: int,
                   homography_interval: int = 1):
        """All kernels for one chunk, nothing copied to the host.  Returns (kp, fit, h_index, attempted, proj)."""
        e = self.engine
        kp = e.decode(heatmaps, width, height, self.keypoint_conf)
        if self.synthesis:
            e.synthesize(kp)
        fit = e.fit(kp, mode=self.fit_mode, K=self.max_iters, thr=self.thr)
        h_index, attempted = e.select(fit.status, homography_interval)
        proj = e.project(fit.H, foot, count, width, height, h_index=h_index)
        return kp, fit, h_index, attempted, proj

    def run(self, heatmaps: torch.Tensor, objects_per_frame: Sequence[dict], width: int, height: int, fps: int,
            homography_interval: int = 1, first_index: int = 0) -> dict:
        """heatmaps (F,57,h,w) float32 CUDA tensor + detect_objects() dicts -> reference result dict."""
        objects_per_frame = list(objects_per_frame)
        foot_h, count_h = objects_to_arrays(objects_per_frame, max(1, max_objects(objects_per_frame)))
        dev = self.engine.device
        foot = torch.from_numpy(foot_h).to(dev, non_blocking=True)
        count = torch.from_numpy(count_h).to(dev, non_blocking=True)
        kp, fit, h_index, attempted, proj = self.run_device(heatmaps, foot, count, width, height, homography_interval)
        from .sharding import download
        return assemble_frames(objects_per_frame, fps, first_index, *download([kp.xy, kp.order, kp.count, fit.used_mask, fit.inlier_mask,
                                                                               fit.status, attempted, h_index, proj.coords_i,
                                                                               proj.in_bounds, proj.bounds]))


def upload_threads(cores: int) -> int:
    return max(2, min(8, cores - 1))


class CoordinateModel:
    """Reference-shaped front end (constructor kwargs and method names of
    eagle/models/coordinate_model.py:49,188,480,557)."""

    def __init__(self, keypoint_conf: float = 0.3, detector_conf: float = 0.35, *, keypoint_model: Callable | None = None,
                 detect_objects: Callable | None = None, device="cuda:0", chunk: int = 128):
        self.device = torch.device(device)
        self.keypoint_conf = keypoint_conf
        self.detector_conf = detector_conf
        self.keypoint_model = keypoint_model
        self._detect_objects = detect_objects
        self.chunk = chunk
        self.path = GeometryPath(device, keypoint_conf)
        self._stream, self._stream_key = None, None
        self.network_batch = BATCH     # frames per keypoint_model call (the reference feeds the network 4 at a time, :20)
        # worker threads of egl_upload_frames: a single core moves 6-10 GB/s into the page-locked rings, the PCIe link 55;
        # measured on a 16-core box: 4 threads 38 GB/s, 6 -> 48, 8 -> 50, 16 -> 44 (they crowd out the assembler and the
        # thread that enqueues the kernels), so: the cores this process may run on minus one, at most 8
        self.copy_threads = upload_threads(len(os.sched_getaffinity(0)))
        self.uploads_in_flight = 1     # 2: chunk c+1's upload starts before chunk c's has been waited for (see DenseStream)
        self.profile = False           # time H2D / kernels / D2H of every chunk with CUDA events (last_stats)
        self.always_propagate = False  # route every clip through PropagatedPath (tests)
        self.piece_frames = 2048       # sparse cadence: frames resident in HBM at a time (12.7 GB of 1080p frames + 5.6 GB of pyramids)
        self.last_stats = {}

    def detect_objects(self, frame: np.ndarray) -> dict:
        if self._detect_objects is None:
            raise RuntimeError("no detector attached: pass detect_objects=<callable frame -> boxes dict>")
        return self._detect_objects(frame)

    @torch.no_grad()
    def _heatmaps_of(self, x: torch.Tensor) -> torch.Tensor:
        """The attached network on a preprocessed (n, 3, 540, 960) tensor, ``network_batch`` frames per call."""
        if self.keypoint_model is None:
            raise RuntimeError("no keypoint network attached: pass keypoint_model=<callable tensor -> heatmaps>")
        b = max(1, int(self.network_batch))
        outs = [self.keypoint_model(x[i:i + b]) for i in range(0, x.shape[0], b)]
        hm = outs[0] if len(outs) == 1 else torch.cat(outs)
        return hm.to(torch.float32).contiguous()

    @torch.no_grad()
    def _heatmaps_dev(self, dev_frames: torch.Tensor) -> torch.Tensor:
        """K1 + the attached network on frames that already are in HBM ((n, H, W, 3) uint8)."""
        return self._heatmaps_of(self.path.engine.preprocess(dev_frames.contiguous()))

    def _upload(self, frames: Sequence[np.ndarray], dst: torch.Tensor) -> None:
        """Host frames -> dst (n, H, W, 3) on the device (egl_upload_frames: worker threads, page-locked rings, H2D per
        slice).  Synchronous; whatever the current stream still does with dst is waited for first."""
        part = [fr if fr.flags["C_CONTIGUOUS"] else np.ascontiguousarray(fr) for fr in frames]
        torch.cuda.current_stream(self.device).synchronize()
        self.path.engine.upload_frames(part, dst, threads=self.copy_threads)

    @torch.no_grad()
    def _heatmaps(self, frames: Sequence[np.ndarray]) -> torch.Tensor:
        host = torch.from_numpy(np.ascontiguousarray(np.stack(frames)))
        return self._heatmaps_dev(host.pin_memory().to(self.device, non_blocking=True))

    @torch.no_grad()
    def detect_keypoints(self, frame: np.ndarray) -> dict:
        """coordinate_model.py:480-518: {label: (xi, yi)} for one BGR frame."""
        h, w = frame.shape[:2]
        kp = self.path.engine.decode(self._heatmaps([frame]), w, h, self.keypoint_conf)
        n = int(kp.count[0, 0])
        order = kp.order[0, :n].cpu().numpy()
        xy = kp.xy[0].cpu().numpy()
        return {LANDMARK_NAMES[int(c)]: (int(xy[c, 0]), int(xy[c, 1])) for c in order}

    def get_coordinates(self, frames, fps: int, num_homography: int = 1, num_keypoint_detection: int = 1,
                        verbose: bool = True, calibration: bool = False) -> dict:
        """coordinate_model.py:188-417.

        With a keypoint interval of 1 the clip streams through eagle_b200.streaming.DenseStream: page-locked staging
        filled by a thread pool, H2D, kernels and the packed D2H on three streams, dict assembly (C extension) in a worker
        thread, the homography cadence carried from chunk to chunk in device memory.  ``last_stats`` has the host-side
        split (detector, staging, assembly seconds and the bytes moved)."""
        if len(frames) == 0:
            return {}
        homography_interval = max(1, int(fps / max(1, num_homography)))
        keypoint_interval = max(1, int(fps / max(1, num_keypoint_detection)))
        if keypoint_interval != 1 or calibration or self.always_propagate:
            return self._get_coordinates_propagated(frames, fps, homography_interval, keypoint_interval, calibration)
        height, width = frames[0].shape[:2]
        # chunks hold whole cadence segments, so the cadence of a chunk needs one number from its predecessor
        chunk = max(homography_interval, self.chunk // homography_interval * homography_interval)
        key = (height, width, chunk)
        if self._stream is None or self._stream_key != key:
            if self._stream is not None:
                self._stream.close()
            from .streaming import DenseStream
            self._stream, self._stream_key = DenseStream(self.path.engine, height, width, chunk, copy_threads=self.copy_threads,
                                                         uploads_in_flight=self.uploads_in_flight), key

        def assemble(objs, first, a, out):
            assemble_frames(objs, fps, first, a["xy"], a["order"], a["count"], None, a["inlier_mask"], a["status"], a["attempted"],
                            a["h_index"], a["coords_i"], a["in_bounds"], a["bounds"], out=out)

        stats = {}
        with torch.no_grad():
            res, all_obj, aborted = self._stream.run(frames, self.detect_objects, self._heatmaps_of, fps, homography_interval,
                                                     self.keypoint_conf, assemble, stats=stats, profile=self.profile)
        self.last_stats = stats
        if aborted:
            # a frame decoded < 4 landmarks: the reference brings optical flow in (:287-311)
            return self._get_coordinates_propagated(frames, fps, homography_interval, keypoint_interval, calibration, all_obj)
        return res

    def _get_coordinates_propagated(self, frames, fps: int, homography_interval: int, keypoint_interval: int, calibration: bool,
                                    all_obj=None) -> dict:
        """Any cadence: frames to HBM, network on the chain heads, PropagatedPath, projection, dict assembly.
        Clips longer than ``piece_frames`` go through in pieces (whole chains each; the boundary state is carried)."""
        from .propagation import FirstPieceTooShort, PropagatedPath, finalize
        e = self.path.engine
        height, width = frames[0].shape[:2]
        if all_obj is None:
            all_obj = [self.detect_objects(f) for f in frames]
        k = keypoint_interval
        piece = max(k, self.piece_frames // k * k)
        prop = PropagatedPath(e, self.keypoint_conf)
        pieces, carry, prev_last = [], None, None
        stats = {"fallback_frames": 0, "repaired_chains": 0, "first_frame_rescue": False, "pieces": 0}
        g0 = 0
        first_piece = piece
        while g0 < len(frames):
            n = min(first_piece if g0 == 0 else piece, len(frames) - g0)
            halo = 0 if carry is None else 1
            buf = torch.empty((halo + n, height, width, 3), dtype=torch.uint8, device=self.device)
            if halo:
                buf[0].copy_(prev_last)
            self._upload(frames[g0:g0 + n], buf[halo:])
            heads = [halo + i for i in range(0, n, k)]
            hm = torch.cat([self._heatmaps_dev(buf[heads[s:s + self.chunk]]) for s in range(0, len(heads), self.chunk)])
            try:
                pieces.append(prop.run(buf, hm, lambda i, buf=buf, g0=g0, halo=halo: self._heatmaps_dev(buf[halo + i - g0:halo + i - g0 + 1]),
                                       k, homography_interval, calibration, first_frame=g0, carry=carry, clip_continues=g0 + n < len(frames)))
            except FirstPieceTooShort:
                first_piece *= 2  # the reference's scan for the first usable frame (:290-297) runs past this piece: take a longer one
                continue
            carry, prev_last = prop.carry_out, buf[-1].clone()
            g0 += n
            for key in ("fallback_frames", "repaired_chains"):
                stats[key] += prop.stats[key]
            stats["first_frame_rescue"] |= prop.stats["first_frame_rescue"]
            stats["pieces"] += 1
        out = finalize(e, pieces)
        self.last_stats = stats
        foot_h, count_h = objects_to_arrays(all_obj, max(1, max_objects(all_obj)))
        proj = e.project(out["H"], torch.from_numpy(foot_h).to(self.device), torch.from_numpy(count_h).to(self.device), width, height,
                         h_index=out["h_index"])
        from .sharding import download
        xy_h, order_h, count_h, hi_h, ci_h, ib_h, bd_h, src_h = download([out["xy"], out["order"], out["count"], out["h_index"], proj.coords_i,
                                                                          proj.in_bounds, proj.bounds, out["src"]])
        return assemble_frames(all_obj, fps, 0, xy_h, order_h, count_h, None, None, None, None, hi_h, ci_h, ib_h, bd_h, kp_src=src_h)
