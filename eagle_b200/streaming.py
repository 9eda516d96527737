"""The per-frame cadence of get_coordinates as a stream: host frames in, reference-format dict out.

coordinate_model.py:188-417 walks the clip frame by frame; with a keypoint interval of 1 every frame is independent up
to the homography cadence, so the clip goes through in chunks of whole cadence segments and four things overlap:

    host      detect_objects() of chunk c+1 while egl_upload_frames moves the frames of chunk c+1 to the device (worker
              threads of the library: 4 MiB slices through a small page-locked ring, H2D per slice on their own streams)
    copy-in   foot points of chunk c+1                                                          [stream copy_in]
    kernels   K1 preprocess -> keypoint network -> K2 decode -> F1 synthesis -> K3 fit ->
              cadence (egl_select_homography_chunk, state carried in device memory) -> K4      [stream compute]
    copy-out  ONE packed record per chunk (about 1 KB per frame) into page-locked memory       [stream copy_out]
    host      dict assembly of chunk c-1 (C extension) in a worker thread

Nothing is read back between chunks; the host blocks only on buffer reuse.  Device memory is two chunks of everything,
whatever the clip length (a 135 k-frame match streams through the same buffers as a 2 k-frame clip).
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Sequence

import numpy as np
import torch

from . import _native as N
from .boxes import max_objects, objects_to_arrays
from .engine import FitResult, GeometryEngine
from .sharding import pack_results


def record_layout(P: int):
    """(name, dtype, per-frame shape) of the packed per-frame record a chunk sends to the host."""
    return (("xy", np.int32, (N.NUM_LANDMARKS, 2)), ("order", np.uint8, (N.ORDER_STRIDE,)), ("count", np.int32, (2,)),
            ("inlier_mask", np.int64, ()), ("status", np.int32, ()), ("attempted", np.uint8, ()), ("h_index", np.int32, ()),
            ("coords_i", np.int64, (P, 2)), ("in_bounds", np.uint8, (P,)), ("bounds", np.float64, (4,)))


def record_bytes(P: int) -> int:
    return sum(int(np.dtype(dt).itemsize * int(np.prod(shape, dtype=np.int64))) for _, dt, shape in record_layout(P))


def unpack_record(buf: np.ndarray, n: int, P: int) -> dict:
    """buf: (n, record_bytes(P)) uint8 on the host -> dict of per-frame arrays (copies, C-contiguous)."""
    out, o = {}, 0
    for name, dt, shape in record_layout(P):
        per = int(np.dtype(dt).itemsize * int(np.prod(shape, dtype=np.int64)))
        out[name] = np.ascontiguousarray(buf[:n, o:o + per]).view(dt).reshape((n,) + tuple(shape))
        o += per
    return out


class DenseStream:
    """Reusable buffers + streams of the per-frame cadence for one frame size."""

    def __init__(self, engine: GeometryEngine, height: int, width: int, chunk: int, copy_threads: int = 8, out_slots: int = 4,
                 uploads_in_flight: int = 1):
        self.e = engine
        dev = engine.device
        self.H, self.W, self.CH = height, width, chunk
        self.copy_in = torch.cuda.Stream(dev)
        self.compute = torch.cuda.Stream(dev)
        self.copy_out = torch.cuda.Stream(dev)
        self.copy_threads = max(1, copy_threads)
        # 2: the upload of chunk c+1 is started before the one of chunk c has been waited for (the library serves two calls
        # at once from separate rings), so the link stays busy while a call drains its last DMAs
        self.uploads_in_flight = 2 if uploads_in_flight >= 2 else 1
        self.pool = ThreadPoolExecutor(max_workers=2)   # runs the (GIL-free) upload calls next to detect_objects
        self.P_cap = 0
        self.slots = []
        for _ in range(2):
            self.slots.append(dict(
                d_frames=torch.empty((chunk, height, width, 3), dtype=torch.uint8, device=dev),
                x=torch.empty((chunk, 3, N.MODEL_H, N.MODEL_W), dtype=torch.float32, device=dev),
                kp=engine.alloc_keypoints(chunk), status=torch.zeros(chunk, dtype=torch.int32, device=dev),
                used=torch.zeros(chunk, dtype=torch.int64, device=dev), inl=torch.zeros(chunk, dtype=torch.int64, device=dev),
                info=torch.zeros((chunk, 4), dtype=torch.int32, device=dev),
                h_index=torch.empty(chunk, dtype=torch.int32, device=dev), attempted=torch.empty(chunk, dtype=torch.uint8, device=dev),
                uploaded=None, computed=None))
        self.out_slots = [dict(buf=None, free=threading.Event()) for _ in range(out_slots)]
        for s in self.out_slots:
            s["free"].set()
        self._grow_points(23)

    def _grow_points(self, P: int) -> None:
        """(Re)allocate everything whose size depends on the largest number of detections per frame."""
        if P <= self.P_cap:
            return
        dev = self.e.device
        torch.cuda.synchronize(dev)
        self.P_cap = P
        for s in self.slots:
            s["h_foot"] = torch.zeros((self.CH, P, 2), dtype=torch.float32, pin_memory=True)
            s["h_cnt"] = torch.zeros((self.CH,), dtype=torch.int32, pin_memory=True)
            s["d_foot"] = torch.zeros((self.CH, P, 2), dtype=torch.float32, device=dev)
            s["d_cnt"] = torch.zeros((self.CH,), dtype=torch.int32, device=dev)
            s["proj"] = self.e.alloc_projection(self.CH, P)
        for s in self.out_slots:
            s["free"].wait()
            s["buf"] = torch.empty((self.CH, record_bytes(P)), dtype=torch.uint8, pin_memory=True)

    def close(self) -> None:
        self.pool.shutdown(wait=True)

    def _upload(self, part, d_frames) -> float:
        import time
        t0 = time.perf_counter()
        self.e.upload_frames(part, d_frames, threads=self.copy_threads)
        return time.perf_counter() - t0

    # ---------------------------------------------------------------------------------------------
    def run(self, frames: Sequence[np.ndarray], detect_objects: Callable, heatmaps_of: Callable[[torch.Tensor], torch.Tensor],
            fps: int, homography_interval: int, keypoint_conf: float, assemble: Callable, few_landmarks_abort: bool = True,
            stats: dict | None = None, profile: bool = False):
        """Returns (result dict, all detections, aborted).  ``aborted`` is True when some frame decoded fewer than four
        landmarks (the reference then brings optical flow in, :287-311): the caller reroutes the clip; the dict is partial.

        heatmaps_of(x) maps the preprocessed (n, 3, 540, 960) float32 tensor of a chunk to (n, 57, h, w) heatmaps on the
        compute stream; assemble(objects, first_index, arrays, out_dict) appends the chunk's frames to the result."""
        import time
        e, dev = self.e, self.e.device
        F = len(frames)
        CH, Hh, Ww = self.CH, self.H, self.W
        assert CH % homography_interval == 0 or CH >= F, "chunks must hold whole cadence segments"
        H_all = torch.zeros((F, 9), dtype=torch.float64, device=dev)
        carry = torch.full((1,), -1, dtype=torch.int32, device=dev)
        few = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.compute.wait_stream(torch.cuda.current_stream(dev))
        res: dict = {}
        all_obj: list = []
        work: queue.Queue = queue.Queue()
        failure: list = []
        t_asm = [0.0]

        def assembler():
            while True:
                item = work.get()
                if item is None:
                    return
                first, n, P, objs, oslot, done = item
                try:
                    done.synchronize()
                    t0 = time.perf_counter()
                    arrays = unpack_record(oslot["buf"].numpy()[:, :record_bytes(P)], n, P)
                    oslot["free"].set()
                    if not failure:
                        assemble(objs, first, arrays, res)
                    t_asm[0] += time.perf_counter() - t0
                except BaseException as exc:  # surfaced by the caller after the join
                    failure.append(exc)
                    oslot["free"].set()

        worker = threading.Thread(target=assembler, name="eagle-assemble", daemon=True)
        worker.start()
        t_host = {"detect": 0.0, "stage": 0.0, "upload": 0.0}
        # frames that already live in page-locked memory (one (F,H,W,3) uint8 torch tensor) are copied from where they are
        pinned_src = torch.is_tensor(frames) and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_pinned()
        frames_np = frames.numpy() if pinned_src else frames
        marks = []  # profile: (h2d start, h2d end, kernels start, kernels end, d2h start, d2h end) events per chunk
        ev = (lambda st: st.record_event(torch.cuda.Event(enable_timing=True))) if profile else (lambda st: None)

        def start_upload(c):
            """Frames of chunk c -> the device buffer of slot c & 1 (after the kernels that last read it), on the helper thread."""
            if pinned_src:
                return None
            first = c * CH
            s = self.slots[c & 1]
            if s["computed"] is not None:
                s["computed"].synchronize()
            part = [fr if fr.flags["C_CONTIGUOUS"] else np.ascontiguousarray(fr) for fr in frames[first:first + CH]]
            return self.pool.submit(self._upload, part, s["d_frames"])

        futs: dict = {}

        def ensure_upload(c):
            if c not in futs and c * CH < F:
                futs[c] = start_upload(c)

        try:
            for c, first in enumerate(range(0, F, CH)):
                n = min(CH, F - first)
                s = self.slots[c & 1]
                # ---- host: the upload runs in the library's worker threads while the detector works on the same frames
                t0 = time.perf_counter()
                ensure_upload(c)
                objs = [detect_objects(frames_np[first + j]) for j in range(n)]
                t1 = time.perf_counter()
                all_obj.extend(objs)
                if self.uploads_in_flight == 2:
                    ensure_upload(c + 1)     # its slot was last read by the kernels of chunk c-1 (start_upload waits for them)
                P = max(1, max_objects(objs))
                if P > self.P_cap:
                    for f_ in list(futs.values()):
                        if f_ is not None:
                            f_.result()
                    self._grow_points(P)
                if s["uploaded"] is not None:
                    s["uploaded"].synchronize()          # the H2D that last read this slot's foot-point staging has finished
                objects_to_arrays(objs, self.P_cap, out=(s["h_foot"].numpy()[:n], s["h_cnt"].numpy()[:n]))
                fut = futs.pop(c)
                if fut is not None:
                    t_host["upload"] += fut.result()
                # the next chunk's upload starts before this chunk's kernels are enqueued (about a millisecond of launches)
                ensure_upload(c + 1)
                t2 = time.perf_counter()
                t_host["detect"] += t1 - t0
                t_host["stage"] += t2 - t1
                P = self.P_cap
                # ---- copy-in
                with torch.cuda.stream(self.copy_in):
                    if pinned_src and s["computed"] is not None:
                        self.copy_in.wait_event(s["computed"])   # the kernels that last read the device buffers have finished
                    m0 = ev(self.copy_in)
                    if pinned_src:
                        s["d_frames"][:n].copy_(frames[first:first + n], non_blocking=True)
                    s["d_foot"][:n].copy_(s["h_foot"][:n], non_blocking=True)
                    s["d_cnt"][:n].copy_(s["h_cnt"][:n], non_blocking=True)
                    m1 = ev(self.copy_in)
                    s["uploaded"] = self.copy_in.record_event()
                # ---- kernels
                with torch.cuda.stream(self.compute):
                    self.compute.wait_event(s["uploaded"])
                    m2 = ev(self.compute)
                    x = e.preprocess(s["d_frames"][:n], out=s["x"][:n])
                    hm = heatmaps_of(x)
                    kp = e.decode(hm, Ww, Hh, keypoint_conf, out=_kp_view(s["kp"], n))
                    e.synthesize(kp)
                    fit = FitResult(H_all[first:first + n], s["used"][:n], s["inl"][:n], s["status"][:n], s["info"][:n])
                    e.fit(kp, out=fit)
                    few |= (kp.count[:, 1] < 4).any().to(torch.int32)
                    h_index, attempted = e.select_chunk(fit.status, homography_interval, first, carry, s["h_index"][:n], s["attempted"][:n])
                    proj = _proj_view(s["proj"], n)
                    e.project(H_all, s["d_foot"][:n], s["d_cnt"][:n], Ww, Hh, h_index=h_index, out=proj)
                    rec = pack_results([kp.xy, kp.order, kp.count, fit.inlier_mask, fit.status, attempted, h_index, proj.coords_i,
                                        proj.in_bounds, proj.bounds])
                    m3 = ev(self.compute)
                    s["computed"] = self.compute.record_event()
                # ---- copy-out
                oslot = self.out_slots[c % len(self.out_slots)]
                oslot["free"].wait()
                oslot["free"].clear()
                with torch.cuda.stream(self.copy_out):
                    self.copy_out.wait_event(s["computed"])
                    m4 = ev(self.copy_out)
                    oslot["buf"][:n, :rec.shape[1]].copy_(rec, non_blocking=True)
                    rec.record_stream(self.copy_out)
                    m5 = ev(self.copy_out)
                    done = self.copy_out.record_event()
                if profile:
                    marks.append((m0, m1, m2, m3, m4, m5))
                work.put((first, n, P, objs, oslot, done))
        finally:
            for f_ in futs.values():
                if f_ is not None and not f_.done():
                    try:
                        f_.result()    # an upload still in flight writes into this object's buffers
                    except Exception:  # noqa: BLE001
                        pass
            work.put(None)
            worker.join()
        torch.cuda.current_stream(dev).wait_stream(self.compute)
        if failure:
            raise failure[0]
        aborted = bool(few.item()) and few_landmarks_abort
        if stats is not None and profile:
            torch.cuda.synchronize(dev)
            stats.update(h2d_ms=sum(a.elapsed_time(b) for a, b, *_ in marks), kernels_ms=sum(m[2].elapsed_time(m[3]) for m in marks),
                         d2h_ms=sum(m[4].elapsed_time(m[5]) for m in marks))
        if stats is not None:
            stats.update(detect_s=t_host["detect"], stage_s=t_host["stage"], upload_s=t_host["upload"], assemble_s=t_asm[0], chunks=(F + CH - 1) // CH,
                         chunk_frames=CH, h2d_bytes=F * (Hh * Ww * 3 + self.P_cap * 8 + 4), d2h_bytes=F * record_bytes(self.P_cap))
        return res, all_obj, aborted


def _kp_view(kp, n: int):
    from .engine import KeypointSet
    return KeypointSet(kp.flat[:n], kp.score[:n], kp.xy[:n], kp.order[:n], kp.count[:n])


def _proj_view(pr, n: int):
    from .engine import Projection
    return Projection(pr.coords[:n], pr.coords_i[:n], pr.in_bounds[:n], pr.bounds[:n])
