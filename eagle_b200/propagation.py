"""Sparse keypoint cadence on the GPU: network keypoints every ``keypoint_interval`` frames, Lucas-Kanade
propagation in between (coordinate_model.py:206, 277-367, 419-478, 520-555).

How the reference's frame loop is laid out here
  * A *chain* is one network frame (the "head", i % keypoint_interval == 0) plus the frames up to the
    next head.  Inside a chain every frame depends on the one before (flow from its final keypoints);
    chains are independent of each other unless a head decodes fewer than four landmarks (:287-311)
    or a failed fit leaves the retry flag set across the chain boundary (:333,:350-367).
  * All per-frame state lives in HBM as (step, chain, ...) arrays, so "advance every chain by one
    frame" is one launch of each kernel over a contiguous slice: track -> filter -> synthesise ->
    [calibrate] -> fit (only where the cadence asks) -> commit.  A clip needs keypoint_interval such
    rounds, whatever its length.
  * That parallel pass speculates "heads are independent, no retry flag crosses a boundary".  The few
    chains for which that turns out wrong are re-run one by one, in order, with the same kernels on
    single-chain slices -- the reference's sequential semantics, exactly.
  * The host reads back once per clip (head keypoint counts, retry flags, per-frame flow counts); a chain in
    which some frame's flow kept fewer than four points -- the reference then asks the network for that
    frame, :316-320 -- is one of those re-run chains.

Nothing here computes on the CPU; the arithmetic is in csrc/flow.cu, fit.cu, synthesize.cu.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from . import _native as N
from .engine import FitResult, GeometryEngine, KeypointSet
from .pitch import NUM_LANDMARKS

LK_MAX_LEVEL, LK_MAX_COUNT, LK_EPS = 2, 10, 0.03  # lk_params, coordinate_model.py:65


class FirstPieceTooShort(RuntimeError):
    """Frame 0 decoded < 4 landmarks and no frame of the first piece has >= 4: the reference's forward scan
    (coordinate_model.py:290-297) would go on into later frames, so the caller must retry with a longer first piece."""


class _Sets:
    """(step, chain) arrays of keypoint sets and fit results."""

    def __init__(self, k: int, nc: int, dev):
        z = lambda *shape, dtype: torch.zeros(shape, dtype=dtype, device=dev)
        self.xy = z(k, nc, NUM_LANDMARKS, 2, dtype=torch.int32)
        self.order = z(k, nc, N.ORDER_STRIDE, dtype=torch.uint8)
        self.count = z(k, nc, 2, dtype=torch.int32)
        self.src = z(k, nc, N.ORDER_STRIDE, dtype=torch.uint8)
        self.H = z(k, nc, 9, dtype=torch.float64)
        self.used = z(k, nc, dtype=torch.int64)
        self.inl = z(k, nc, dtype=torch.int64)
        self.status = torch.full((k, nc), N.FIT_SKIPPED, dtype=torch.int32, device=dev)
        self.info = z(k, nc, 4, dtype=torch.int32)
        self.fit_ok = z(k, nc, dtype=torch.uint8)
        self.cal_err = z(k, nc, dtype=torch.int32)
        self.retry = z(nc, dtype=torch.uint8)

    def kp(self, s: int, c0: int, c1: int) -> KeypointSet:
        return KeypointSet(None, None, self.xy[s, c0:c1], self.order[s, c0:c1], self.count[s, c0:c1], self.src[s, c0:c1])

    def fit(self, s: int, c0: int, c1: int) -> FitResult:
        return FitResult(self.H[s, c0:c1], self.used[s, c0:c1], self.inl[s, c0:c1], self.status[s, c0:c1], self.info[s, c0:c1])


def _new_set(n: int, dev) -> KeypointSet:
    return KeypointSet(torch.empty((n, NUM_LANDMARKS), dtype=torch.int32, device=dev), torch.empty((n, NUM_LANDMARKS), dtype=torch.float32, device=dev),
                       torch.zeros((n, NUM_LANDMARKS, 2), dtype=torch.int32, device=dev), torch.zeros((n, N.ORDER_STRIDE), dtype=torch.uint8, device=dev),
                       torch.zeros((n, 2), dtype=torch.int32, device=dev), torch.zeros((n, N.ORDER_STRIDE), dtype=torch.uint8, device=dev))


def _clone(a: KeypointSet) -> KeypointSet:
    return KeypointSet(None, None, a.xy.clone(), a.order.clone(), a.count.clone(), a.src.clone())


def _assign(dst: KeypointSet, src: KeypointSet) -> None:
    dst.xy.copy_(src.xy); dst.order.copy_(src.order); dst.count.copy_(src.count); dst.src.copy_(src.src)


class PropagatedPath:
    """The frame loop of get_coordinates for any keypoint / homography cadence, frames resident in HBM."""

    def __init__(self, engine: GeometryEngine, keypoint_conf: float = 0.3, fit_mode: int = N.FIT_CV2_COMPAT, max_iters: int = 2000,
                 thr: float = 5.0):
        self.e = engine
        self._side = None
        self._pyr_stream = None
        self._rounds = None
        self.keypoint_conf = keypoint_conf
        self.fit_mode, self.max_iters, self.thr = fit_mode, max_iters, thr
        self.stats = {}

    # ---------------------------------------------------------------------------------------------
    def run(self, frames: torch.Tensor, head_heatmaps: torch.Tensor, detect: Callable[[int], torch.Tensor] | None,
            keypoint_interval: int, homography_interval: int, calibration: bool = False, *, first_frame: int = 0,
            carry: dict | None = None, clip_continues: bool = False) -> dict:
        """frames (F, H, W, 3) uint8 BGR CUDA; head_heatmaps (ceil(F/k), 57, h, w) float32 CUDA = the network's
        output for frames 0, k, 2k, ...; detect(i) -> (1, 57, h, w) heatmaps of frame i on demand.

        Long clips go through in pieces: ``first_frame`` (a multiple of k) is the clip index of this piece's
        first frame and ``carry`` what the previous piece left in ``self.carry_out``; frames then has ONE extra
        leading frame, the last frame of the previous piece (the flow into a head with < 4 landmarks starts there).

        Returns device tensors in frame order: xy, order, count, src (the "Keypoints" of every frame),
        H (F, 9), fit_ok (F,), h_index (F,) (valid for a single piece; see ``finalize``) and counters in ``self.stats``."""
        assert (carry is None) == (first_frame == 0), "carry comes with every piece but the first"
        self.start(frames, head_heatmaps, detect, keypoint_interval, homography_interval, calibration, first_frame=first_frame,
                   clip_continues=clip_continues)
        self.repair(carry)
        return self.outputs()

    def start(self, frames: torch.Tensor, head_heatmaps: torch.Tensor, detect, keypoint_interval: int, homography_interval: int,
              calibration: bool = False, *, first_frame: int = 0, clip_continues: bool = False) -> None:
        """The parallel pass of one piece (everything that does not need the previous piece's final state).
        Pieces after the first (first_frame > 0) carry one extra leading frame.  ``repair`` must follow.
        clip_continues: more frames follow this piece (only matters for the first-frame rescue, see FirstPieceTooShort)."""
        # The rounds run on high-priority streams of this object, the pyramid pass on a normal one: their short launches
        # are then scheduled into the machine as thread-block slots free up instead of queueing behind the pyramid's grid.
        dev = frames.device
        caller = torch.cuda.current_stream(dev)
        if self._rounds is None:
            self._rounds = torch.cuda.Stream(dev, priority=-1)
            self._side = torch.cuda.Stream(dev, priority=-1)
            self._pyr_stream = torch.cuda.Stream(dev)
        self._rounds.wait_stream(caller)
        with torch.cuda.stream(self._rounds):
            self._start(frames, head_heatmaps, detect, keypoint_interval, homography_interval, calibration, first_frame, clip_continues)
        caller.wait_stream(self._rounds)
        for t in (frames, head_heatmaps):
            t.record_stream(self._rounds)

    def _start(self, frames, head_heatmaps, detect, keypoint_interval, homography_interval, calibration, first_frame, clip_continues) -> None:
        self.clip_continues = clip_continues
        e = self.e
        k = int(keypoint_interval)
        self.g0 = int(first_frame)
        self.base = 0 if self.g0 == 0 else 1            # device index of this piece's frame 0
        assert self.g0 % k == 0, "pieces start on a chain head"
        F, Himg, Wimg = frames.shape[0] - self.base, frames.shape[1], frames.shape[2]
        nc = (F + k - 1) // k
        dev = frames.device
        assert head_heatmaps.shape[0] == nc, "one heatmap stack per chain head"
        self.frames, self.k, self.F, self.Himg, self.Wimg = frames, k, F, Himg, Wimg
        self.calibration = calibration
        self.detect = detect
        self.carry = None
        # Pyramids one step of every chain at a time, on a stream of their own: the tracker of round s only needs steps s
        # and s + 1, so the rounds (latency-bound launches that leave the machine nearly idle) run underneath the
        # HBM-bound pyramid pass instead of after it.
        main = torch.cuda.current_stream(dev)
        self.pyr = e.alloc_pyramid(frames.shape[0], Himg, Wimg, LK_MAX_LEVEL, device=dev)
        pyr_ready = []
        self._pyr_stream.wait_stream(main)
        with torch.cuda.stream(self._pyr_stream):
            for s_ in range(min(k, F)):
                if s_ == 0 and self.base:
                    e.gray_pyramid_step(frames, self.pyr, 0, frames.shape[0], LK_MAX_LEVEL)    # the carried leading frame
                e.gray_pyramid_step(frames, self.pyr, self.base + s_, k, LK_MAX_LEVEL)
                pyr_ready.append(self._pyr_stream.record_event())
        self.pyr.record_stream(self._pyr_stream)
        frames.record_stream(self._pyr_stream)
        st = self.st = _Sets(k, nc, dev)
        idx = np.arange(k)[:, None] + np.arange(nc)[None, :] * k
        self.sched_h = (((idx + self.g0) % homography_interval) == 0) & (idx < F)
        self.sched = torch.from_numpy(self.sched_h.astype(np.uint8)).to(dev)
        self.extra_mem: dict[int, KeypointSet] = {}     # entries the first-frame rescue wrote into the reference's `mem`
        self.detected: dict[int, KeypointSet] = {}      # cache of on-demand detections (not semantic: `mem[i]` is only read at frame i)
        self.stats = {"fallback_frames": 0, "repaired_chains": 0, "first_frame_rescue": False}

        # heads: decode straight into step 0 of the state, keep a pristine copy (the reference's `mem`)
        tmp = _new_set(nc, dev)
        e.decode(head_heatmaps, Wimg, Himg, self.keypoint_conf, out=KeypointSet(tmp.flat, tmp.score, st.xy[0], st.order[0], st.count[0]))
        self.heads = _clone(st.kp(0, 0, nc))

        # ---- parallel pass: every chain advances one frame per round; nothing is read back in between, so the
        # host runs ahead of the GPU.  (Splitting the chains over 2 / 4 / 8 streams to overlap the short,
        # latency-bound launches of a round was measured slower -- 9.6 / 12.7 / 18.9 ms against 9.45 ms for a
        # 2250-frame clip: the extra host-side launch work costs more than the overlap gains.)
        #
        # Tracking a point does not depend on which other points are in the set, only the filters do.  So the
        # tracker of round s+1 starts on a second stream as soon as round s has its synthesised (and calibrated)
        # keypoints, on a snapshot of that full set, while the main stream fits and commits round s; the filter
        # of round s+1 then picks, by channel, the tracked points of whatever set the commit left.  The two
        # longest latency-bound launches of a round (tracker, refit) overlap instead of queueing.
        flow_cnt = torch.full((k, nc), 1 << 20, dtype=torch.int32, device=dev)
        pts = torch.empty((k, nc, N.ORDER_STRIDE, 2), dtype=torch.float32, device=dev)
        pst = torch.empty((k, nc, N.ORDER_STRIDE), dtype=torch.uint8, device=dev)
        side = self._side
        tracked = None
        for s in range(k):
            n_s = (F - s + k - 1) // k
            if n_s <= 0:
                break
            kp = st.kp(s, 0, n_s)
            if s > 0:
                main.wait_event(tracked)
                e.filter_flow(frames, self.base + s, k, st.kp(s - 1, 0, n_s), pts[s, :n_s], pst[s, :n_s], kp)
                flow_cnt[s, :n_s].copy_(st.count[s, :n_s, 0])
            self._prepare(s, 0, n_s)
            n_next = (F - s - 1 + k - 1) // k
            if s + 1 < k and n_next > 0:
                snap = KeypointSet(None, None, st.xy[s, :n_next], st.order[s, :n_next].clone(), st.count[s, :n_next].clone(), None)
                ready = main.record_event()
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    side.wait_event(pyr_ready[s]); side.wait_event(pyr_ready[s + 1])
                    e.track(self.pyr, Himg, Wimg, snap, self.base + s, self.base + s + 1, k, LK_MAX_LEVEL, LK_MAX_COUNT, LK_EPS,
                            out=(pts[s + 1, :n_next], pst[s + 1, :n_next]))
                    tracked = side.record_event()
                snap.order.record_stream(side); snap.count.record_stream(side)
            self._fit_commit(s, 0, n_s)

        main.wait_stream(self._pyr_stream)   # repairs and single-frame flows may touch any pyramid row
        self._flow_cnt = flow_cnt

    def repair(self, carry: dict | None) -> None:
        """Read back once, then re-run, in frame order, the chains whose speculation was wrong.  ``carry`` =
        ``carry_out`` of the previous piece (None for the first piece of a clip)."""
        st, k, F = self.st, self.k, self.F
        nc = st.retry.numel()
        flow_cnt = self._flow_cnt
        self.carry = carry
        assert (carry is None) == (self.base == 0)
        # ---- one read-back, then the repairs in frame order: chains whose speculation was wrong
        #  * a frame whose flow kept < 4 points (the reference then asks the network, :316-320),
        #  * a head with < 4 landmarks (flow from the previous chain joins in, :308-311; frame 0: :288-307),
        #  * a retry flag that crosses into a head the cadence does not schedule (:333).
        host = torch.stack([self.heads.count[:, 0], st.retry.to(torch.int32), (flow_cnt < 4).any(dim=0).to(torch.int32)]).cpu().numpy()
        head_cnt, retry_final, short = host[0], host[1].copy(), host[2]
        rerun_upto = -1
        if carry is None and head_cnt[0] < 4 and (F > 1 or self.clip_continues):
            rerun_upto = self._rescue_first_frame()
        for c in range(nc):
            incoming = int(retry_final[c - 1]) if c > 0 else (int(carry["retry"]) if carry is not None else 0)
            head = self.extra_mem.get(c * k)
            cnt_c = int(head.count[0, 0]) if head is not None else int(head_cnt[c])
            has_pred = c > 0 or carry is not None
            need = c <= rerun_upto or bool(short[c]) or (has_pred and (cnt_c < 4 or (incoming and not self.sched_h[0, c])))
            if need:
                self._run_chain(c, incoming)
                retry_final[c] = int(st.retry[c].item())
                self.stats["repaired_chains"] += 1
        last = F - 1
        self.carry_out = {"retry": int(retry_final[nc - 1]), "kp": _clone(st.kp(last % k, last // k, last // k + 1))}

    def outputs(self) -> dict:
        e, st, k, F = self.e, self.st, self.k, self.F
        nc = st.retry.numel()
        calibration = self.calibration
        # ---- back to frame order (chain-major == frame order), cadence lookup of the H each frame uses
        fo = lambda t: t.transpose(0, 1).reshape((nc * k,) + tuple(t.shape[2:]))[:F].contiguous()
        if calibration:
            bad = torch.nonzero(fo(st.cal_err))
            if bad.numel() > 0:
                # coordinate_model.py:548 -- grid_hsv[OFFSET, OFFSET] on a block clipped at the top / left edge
                raise IndexError(f"index 3 is out of bounds for axis 0 with size 3 (calibrate_keypoints, frame {int(bad[0, 0])})")
        out = {"xy": fo(st.xy), "order": fo(st.order), "count": fo(st.count), "src": fo(st.src), "H": fo(st.H), "fit_ok": fo(st.fit_ok),
               "status": fo(st.status), "inlier_mask": fo(st.inl)}
        sel_status = torch.where(out["fit_ok"] != 0, torch.full_like(out["status"], N.FIT_OK), torch.full_like(out["status"], N.FIT_NO_MODEL))
        out["h_index"], _ = e.select(sel_status, 1)
        return out

    # ---------------------------------------------------------------------------------------------
    def _flow(self, s: int, c0: int, c1: int) -> None:
        """calculate_optical_flow(frame_i, gray_{i-1}, keypoints_{i-1}, gray_i) for the frames i = c*k + s (:315)."""
        e, st, k = self.e, self.st, self.k
        prev = st.kp(s - 1, c0, c1)
        p0, n0 = self.base + c0 * k + s - 1, self.base + c0 * k + s
        pts, status = e.track(self.pyr, self.Himg, self.Wimg, prev, p0, n0, k, LK_MAX_LEVEL, LK_MAX_COUNT, LK_EPS)
        e.filter_flow(self.frames, n0, k, prev, pts, status, st.kp(s, c0, c1))

    def _flow_single(self, prev: KeypointSet, prev_frame: int, next_frame: int, hue_frame: int) -> KeypointSet:
        e = self.e
        out = _new_set(1, self.frames.device)
        b = self.base  # frame arguments are indices within this piece (-1 = the carried frame)
        pts, status = e.track(self.pyr, self.Himg, self.Wimg, prev, b + prev_frame, b + next_frame, 1, LK_MAX_LEVEL, LK_MAX_COUNT, LK_EPS)
        e.filter_flow(self.frames, b + hue_frame, 1, prev, pts, status, out)
        return out

    def _detect_set(self, i: int) -> KeypointSet:
        """mem.get(i, self.detect_keypoints(frame)) (:318): the network on one frame, decoded."""
        if i in self.extra_mem:
            return self.extra_mem[i]
        if i in self.detected:
            return self.detected[i]
        if i % self.k == 0:
            c = i // self.k
            return KeypointSet(None, None, self.heads.xy[c:c + 1], self.heads.order[c:c + 1], self.heads.count[c:c + 1], self.heads.src[c:c + 1])
        if self.detect is None:
            raise RuntimeError(f"frame {i}: optical flow kept fewer than 4 keypoints and no keypoint network is attached for the "
                               "fallback detection (coordinate_model.py:316-320)")
        hm = self.detect(self.g0 + i)
        d = _new_set(1, self.frames.device)
        self.e.decode(hm, self.Wimg, self.Himg, self.keypoint_conf, out=d)
        self.detected[i] = d
        self.stats["fallback_frames"] += 1
        return d

    def _fallback(self, s: int, c: int) -> None:
        """:316-320,324: keypoints = {**detected, **flowed}, then {**keypoints, **detected}."""
        e, st = self.e, self.st
        d = self._detect_set(c * self.k + s)
        a = _clone(d)
        cur = st.kp(s, c, c + 1)
        e.merge(a, cur)
        e.merge(a, d)
        _assign(cur, a)

    def _finish(self, s: int, c0: int, c1: int) -> None:
        """:326-367 for frames i = c*k + s: synthesis, calibration, fit where the cadence asks, inlier commit."""
        self._prepare(s, c0, c1)
        self._fit_commit(s, c0, c1)

    def _prepare(self, s: int, c0: int, c1: int) -> None:
        """:326-330: synthesis and calibration -- after this the set is what the next frame's flow starts from,
        unless the fit below narrows it to its inliers."""
        e, st, k = self.e, self.st, self.k
        kp = st.kp(s, c0, c1)
        e.synthesize(kp)
        if self.calibration:
            st.cal_err[s, c0:c1].zero_()
            e.calibrate(self.frames, self.base + c0 * k + s, k, kp, st.cal_err[s, c0:c1])

    def _fit_commit(self, s: int, c0: int, c1: int) -> None:
        """:333-367: fit where the cadence asks, inlier commit, retry flag."""
        e, st = self.e, self.st
        kp = st.kp(s, c0, c1)
        fit = st.fit(s, c0, c1)
        sched, retry = self.sched[s, c0:c1], st.retry[c0:c1]
        e.fit(kp, mode=self.fit_mode, K=self.max_iters, thr=self.thr, out=fit, sched=sched, retry=retry)
        e.commit(kp, fit, sched, retry, st.fit_ok[s, c0:c1])

    def _run_chain(self, c: int, incoming_retry: int) -> None:
        """One chain, frame after frame, with the true predecessor state (the reference's loop, literally)."""
        e, st, k, F = self.e, self.st, self.k, self.F
        st.retry[c] = incoming_retry
        i = c * k
        cur = st.kp(0, c, c + 1)
        head = KeypointSet(None, None, self.heads.xy[c:c + 1], self.heads.order[c:c + 1], self.heads.count[c:c + 1], self.heads.src[c:c + 1])
        if c == 0 and self.carry is None:
            _assign(cur, head)                      # :285 keypoints = decoded; the rescue only rewrote mem
            if 0 in self.extra_mem:
                e.merge(cur, self.extra_mem[0])     # :324
        else:
            m = self.extra_mem.get(i, head)         # :285 mem.get(i)
            _assign(cur, m)
            if int(m.count[0, 0]) < 4:              # :308-311
                pred = st.kp(k - 1, c - 1, c) if c > 0 else self.carry["kp"]
                flowed = self._flow_single(pred, i - 1, i, i)
                e.merge(cur, flowed)
                e.merge(cur, m)                     # :324
        self._finish(0, c, c + 1)
        for s in range(1, k):
            i = c * k + s
            if i >= F:
                break
            self._flow(s, c, c + 1)
            if int(st.count[s, c, 0].item()) < 4:
                self._fallback(s, c)
            elif i in self.extra_mem:
                e.merge(st.kp(s, c, c + 1), self.extra_mem[i])  # :322,:324
            self._finish(s, c, c + 1)

    def _rescue_first_frame(self) -> int:
        """:288-307: frame 0 decoded < 4 landmarks -- find the first later frame with >= 4, flow its keypoints
        backwards to frame 0 and merge them into ``mem``.  Returns the last chain that has to be re-run."""
        F, k = self.F, self.k
        self.stats["first_frame_rescue"] = True
        prev = None
        j = 0
        for j in range(1, F):
            d = self._detect_set(j)
            self.extra_mem.setdefault(j, d)
            if int(d.count[0, 0]) >= 4:
                prev = d
                break
        if prev is None:
            if self.clip_continues:
                raise FirstPieceTooShort(f"no frame with >= 4 landmarks among the first {F}")
            return -1
        next_idx = j
        for jj in range(j - 1, -1, -1):
            flowed = self._flow_single(prev, jj, next_idx, jj)  # (prev_frame, prev_gray) = frame jj, curr_gray = frame jj+1 (:303)
            if int(flowed.count[0, 0]) > 0:
                prev = flowed
            base = _clone(prev)
            existing = self.extra_mem.get(jj)
            if existing is None and jj % k == 0:
                c = jj // k
                existing = KeypointSet(None, None, self.heads.xy[c:c + 1], self.heads.order[c:c + 1], self.heads.count[c:c + 1], self.heads.src[c:c + 1])
            if existing is not None:
                self.e.merge(base, existing)
            self.extra_mem[jj] = base
            next_idx = jj
        return j // k


def finalize(engine: GeometryEngine, pieces: list) -> dict:
    """Concatenate the outputs of consecutive ``run`` calls and look up, over the whole clip, the homography each
    frame uses (the last successful fit at or before it, coordinate_model.py:375-378)."""
    out = {key: (torch.cat([p[key] for p in pieces]) if len(pieces) > 1 else pieces[0][key])
           for key in ("xy", "order", "count", "src", "H", "fit_ok", "status", "inlier_mask")}
    sel = torch.where(out["fit_ok"] != 0, torch.full_like(out["status"], N.FIT_OK), torch.full_like(out["status"], N.FIT_NO_MODEL))
    out["h_index"], _ = engine.select(sel, 1)
    return out
