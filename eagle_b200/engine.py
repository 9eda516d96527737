"""Host side of the geometry path: device buffers + the C-ABI calls, one method per kernel group.

PyTorch is used for what it is good at here -- allocating device memory, streams, pinned host
buffers and (in sharding.py) the NCCL plumbing.  All arithmetic happens in libeagle_b200.so; every
method only enqueues kernels on the current CUDA stream and returns device tensors.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _native as N
from .pitch import NUM_LANDMARKS


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require(cond: bool, msg: str) -> None:
    """Argument checks of the host wrapper (kept as real exceptions: asserts vanish under -O)."""
    if not cond:
        raise ValueError(msg)


@dataclass
class KeypointSet:
    """Per-frame landmark detections in HBM (what K2 writes and F1 / K3 read)."""
    flat: torch.Tensor    # (F, 57) int32   argmax flat index per channel
    score: torch.Tensor   # (F, 57) float32 heatmap maximum per channel
    xy: torch.Tensor      # (F, 57, 2) int32 image pixel per channel
    order: torch.Tensor   # (F, 64) uint8   channels kept, in the reference's dict order
    count: torch.Tensor   # (F, 2) int32    {entries in order, of which decoded from heatmaps}
    src: torch.Tensor | None = None  # (F, 64) uint8 KP_* type tag per channel (keypoint propagation only)

    @property
    def n_frames(self) -> int:
        return self.xy.shape[0]


@dataclass
class FitResult:
    H: torch.Tensor            # (F, 9) float64, row-major image->pitch homography
    used_mask: torch.Tensor    # (F,) int64 bit-per-channel: landmarks handed to the fit
    inlier_mask: torch.Tensor  # (F,) int64 bit-per-channel: inliers of the refined H
    status: torch.Tensor       # (F,) int32 FIT_OK / FIT_FEW_POINTS / FIT_NO_MODEL
    info: torch.Tensor         # (F, 4) int32 {points, inliers, winning hypothesis, hypotheses evaluated}


@dataclass
class Projection:
    coords: torch.Tensor    # (F, P, 2) float32 pitch metres before truncation
    coords_i: torch.Tensor  # (F, P, 2) int64   truncated like numpy .astype(int)
    in_bounds: torch.Tensor  # (F, P) uint8
    bounds: torch.Tensor    # (F, 4) float64 [bottom_left, top_left, top_right, bottom_right] x; NaN = None


def letterbox_geometry(h: int, w: int, imgsz: int = 960, stride: int = 32):
    """ultralytics 8.3.184 LetterBox(new_shape=(imgsz, imgsz), auto=True, scaleup=True, center=True, stride=stride) for an
    h x w frame: (new_w, new_h, pad_left, pad_top, out_w, out_h) -- the size cv2.resize is asked for, the border
    cv2.copyMakeBorder adds on the left / top, and the shape of the detector's input."""
    r = min(imgsz / h, imgsz / w)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = (imgsz - new_w) % stride, (imgsz - new_h) % stride   # auto=True: pad only to a multiple of the stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_w, new_h, left, top, new_w + left + right, new_h + top + bottom


class GeometryEngine:
    """Thin, stateless-per-call wrapper; ``device`` must be a CUDA device."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise N.NativeError("GeometryEngine needs a CUDA device; there is no CPU path")

    # -- host frames -> device ---------------------------------------------------------------
    def upload_frames(self, frames, out: torch.Tensor, threads: int = 8) -> torch.Tensor:
        """frames: sequence of n C-contiguous uint8 numpy arrays of one shape (pageable host memory, as the reference's
        frame list holds them); out: (>= n, ...) uint8 CUDA tensor whose rows have that many bytes.  Synchronous."""
        import ctypes as C
        n = len(frames)
        if n == 0:
            return out[:0]
        nbytes = frames[0].nbytes
        _require(out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.shape[0] >= n and out[0].numel() == nbytes,
                 "upload_frames: out must be a contiguous CUDA uint8 tensor with one row per frame")
        ptrs = (C.c_void_p * n)()
        for i, fr in enumerate(frames):
            _require(fr.nbytes == nbytes and fr.flags["C_CONTIGUOUS"] and fr.dtype.itemsize == 1,
                     "upload_frames: frames must be C-contiguous uint8 arrays of one size")
            ptrs[i] = fr.ctypes.data
        with torch.cuda.device(out.device):
            N.check(N.lib.egl_upload_frames(ptrs, n, nbytes, _ptr(out), int(threads)), "egl_upload_frames")
        return out[:n]

    # -- K1 ---------------------------------------------------------------------------------
    def preprocess(self, frames: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """frames (F, H, W, 3) uint8 BGR on the device -> (F, 3, 540, 960) float32."""
        _require(frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_cuda,
                 "preprocess: frames must be a CUDA uint8 tensor of shape (F, H, W, 3)")
        _require(frames.stride(3) == 1 and frames.stride(2) == 3, "preprocess: pixels must be packed B,G,R bytes (strides 3, 1)")
        F, H, W, _ = frames.shape
        if out is None:
            out = torch.empty((F, 3, N.MODEL_H, N.MODEL_W), dtype=torch.float32, device=frames.device)
        with torch.cuda.device(frames.device):
            N.check(N.lib.egl_preprocess_u8(_ptr(frames), F, H, W, frames.stride(1), frames.stride(0) if F > 1 else H * frames.stride(1),
                                            _ptr(out), _stream()), "egl_preprocess_u8")
        return out

    def preprocess_with_detector_input(self, frames: torch.Tensor, out: torch.Tensor | None = None, out_detector: torch.Tensor | None = None,
                                       imgsz: int = 960, stride: int = 32):
        """K1 with its second output: (F, 3, 540, 960) for the keypoint network AND the detector's letterboxed input
        (F, 3, detector_h, 960) float32 RGB in [0, 1] -- what ultralytics' LetterBox + predictor preprocess make of the frame
        before `self.detector_model(frame, ...)` (coordinate_model.py:568) runs -- from ONE read of the uint8 frames.
        Frames whose letterbox resize is not 960x540 (see letterbox_geometry) are rejected."""
        _require(frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_cuda,
                 "preprocess: frames must be a CUDA uint8 tensor of shape (F, H, W, 3)")
        _require(frames.stride(3) == 1 and frames.stride(2) == 3, "preprocess: pixels must be packed B,G,R bytes (strides 3, 1)")
        F, H, W, _ = frames.shape
        new_w, new_h, left, top, det_w, det_h = letterbox_geometry(H, W, imgsz, stride)
        _require((new_w, new_h) == (N.MODEL_W, N.MODEL_H) and left == 0 and det_w == N.MODEL_W,
                 f"preprocess_with_detector_input: a {W}x{H} frame letterboxes to {new_w}x{new_h} at imgsz {imgsz}, not to 960x540")
        if out is None:
            out = torch.empty((F, 3, N.MODEL_H, N.MODEL_W), dtype=torch.float32, device=frames.device)
        if out_detector is None:
            out_detector = torch.empty((F, 3, det_h, det_w), dtype=torch.float32, device=frames.device)
        _require(tuple(out_detector.shape) == (F, 3, det_h, det_w) and out_detector.is_contiguous() and out_detector.dtype == torch.float32,
                 f"preprocess_with_detector_input: out_detector must be a contiguous float32 tensor of shape {(F, 3, det_h, det_w)}")
        with torch.cuda.device(frames.device):
            N.check(N.lib.egl_preprocess_u8_letterbox(_ptr(frames), F, H, W, frames.stride(1), frames.stride(0) if F > 1 else H * frames.stride(1),
                                                      _ptr(out), _ptr(out_detector), det_h, top, _stream()), "egl_preprocess_u8_letterbox")
        return out, out_detector

    # -- K2 ---------------------------------------------------------------------------------
    def alloc_keypoints(self, F: int) -> KeypointSet:
        dev = self.device
        return KeypointSet(torch.empty((F, NUM_LANDMARKS), dtype=torch.int32, device=dev),
                           torch.empty((F, NUM_LANDMARKS), dtype=torch.float32, device=dev),
                           torch.empty((F, NUM_LANDMARKS, 2), dtype=torch.int32, device=dev),
                           torch.empty((F, N.ORDER_STRIDE), dtype=torch.uint8, device=dev),
                           torch.empty((F, 2), dtype=torch.int32, device=dev))

    def decode(self, heatmaps: torch.Tensor, img_w: int, img_h: int, keypoint_conf: float = 0.3,
               out: KeypointSet | None = None, from_logits: bool = False) -> KeypointSet:
        """heatmaps (F, 57, h, w) float32 contiguous on the device; from_logits=True takes the network's
        pre-sigmoid output and fuses the sigmoid into the arg-max kernel."""
        _require(heatmaps.dtype == torch.float32 and heatmaps.is_cuda and heatmaps.is_contiguous() and heatmaps.dim() == 4,
                 "decode: heatmaps must be a contiguous CUDA float32 tensor of shape (F, 57, h, w)")
        F, C, h, w = heatmaps.shape
        _require(C == NUM_LANDMARKS, f"decode: expected {NUM_LANDMARKS} heatmap channels, got {C}")
        kp = out if out is not None else self.alloc_keypoints(F)
        with torch.cuda.device(heatmaps.device):
            fn = N.lib.egl_decode_logits if from_logits else N.lib.egl_decode_heatmaps
            N.check(fn(_ptr(heatmaps), F, h, w, img_w, img_h, float(keypoint_conf), _ptr(kp.flat), _ptr(kp.score), _ptr(kp.xy),
                       _ptr(kp.order), _ptr(kp.count), _stream()), "egl_decode_logits" if from_logits else "egl_decode_heatmaps")
        return kp

    def refine(self, heatmaps: torch.Tensor, kp: KeypointSet, img_w: int, img_h: int) -> torch.Tensor:
        """Sub-pixel positions (F, 57, 2) float32 of every channel's arg-max (extension; see the header)."""
        F, C, h, w = heatmaps.shape
        _require(heatmaps.dtype == torch.float32 and heatmaps.is_cuda and heatmaps.is_contiguous() and C == NUM_LANDMARKS,
                 "refine: heatmaps must be a contiguous CUDA float32 tensor of shape (F, 57, h, w)")
        sub = torch.empty((F, NUM_LANDMARKS, 2), dtype=torch.float32, device=heatmaps.device)
        with torch.cuda.device(heatmaps.device):
            N.check(N.lib.egl_refine_keypoints(_ptr(heatmaps), F, h, w, img_w, img_h, _ptr(kp.flat), _ptr(sub), _stream()),
                    "egl_refine_keypoints")
        return sub

    # -- F1 ---------------------------------------------------------------------------------
    def synthesize(self, kp: KeypointSet, max_new: int = 30) -> KeypointSet:
        with torch.cuda.device(kp.xy.device):
            N.check(N.lib.egl_synthesize_keypoints(_ptr(kp.xy), _ptr(kp.order), _ptr(kp.count), kp.n_frames, max_new, _stream()),
                    "egl_synthesize_keypoints")
        return kp

    # -- K3 ---------------------------------------------------------------------------------
    def alloc_fit(self, F: int) -> FitResult:
        dev = self.device
        return FitResult(torch.zeros((F, 9), dtype=torch.float64, device=dev), torch.zeros(F, dtype=torch.int64, device=dev),
                         torch.zeros(F, dtype=torch.int64, device=dev), torch.zeros(F, dtype=torch.int32, device=dev),
                         torch.zeros((F, 4), dtype=torch.int32, device=dev))

    def fit(self, kp: KeypointSet, mode: int = N.FIT_CV2_COMPAT, K: int = 2000, hyp: torch.Tensor | None = None,
            seed: int = 0, thr: float = 5.0, confidence: float = 0.995, out: FitResult | None = None,
            sched: torch.Tensor | None = None, retry: torch.Tensor | None = None, sub: torch.Tensor | None = None) -> FitResult:
        """sched / retry (F,) uint8: fit only the frames with sched | retry (the reference's cadence test, :333).
        sub (F, 57, 2) float32: fit on these sub-pixel positions instead of the integer kp.xy (extension)."""
        F = kp.n_frames
        r = out if out is not None else self.alloc_fit(F)
        if hyp is not None:
            _require(hyp.dtype == torch.uint8 and hyp.is_cuda and hyp.is_contiguous() and tuple(hyp.shape) == (F, K, 4),
                     "fit: hyp must be a contiguous CUDA uint8 tensor of shape (F, K, 4)")
        with torch.cuda.device(kp.xy.device):
            if sub is not None:
                _require(sched is None, "fit: sub-pixel positions and a cadence mask cannot be combined")
                _require(sub.dtype == torch.float32 and sub.is_contiguous() and tuple(sub.shape) == (F, NUM_LANDMARKS, 2),
                         "fit: sub must be a contiguous (F, 57, 2) float32 tensor")
                N.check(N.lib.egl_fit_homography_subpixel(_ptr(sub), _ptr(kp.xy), _ptr(kp.order), _ptr(kp.count), F, mode, K, _ptr(hyp),
                                                          seed, float(thr), float(confidence), _ptr(r.H), _ptr(r.used_mask),
                                                          _ptr(r.inlier_mask), _ptr(r.status), _ptr(r.info), _stream()),
                        "egl_fit_homography_subpixel")
            elif sched is None:
                N.check(N.lib.egl_fit_homography(_ptr(kp.xy), _ptr(kp.order), _ptr(kp.count), F, mode, K, _ptr(hyp), seed, float(thr),
                                                 float(confidence), _ptr(r.H), _ptr(r.used_mask), _ptr(r.inlier_mask), _ptr(r.status),
                                                 _ptr(r.info), _stream()), "egl_fit_homography")
            else:
                _require(sched.dtype == torch.uint8 and sched.numel() == F and sched.is_contiguous(), "fit: sched must be (F,) uint8")
                _require(retry is None or (retry.dtype == torch.uint8 and retry.numel() == F and retry.is_contiguous()),
                         "fit: retry must be (F,) uint8")
                N.check(N.lib.egl_fit_homography_masked(_ptr(kp.xy), _ptr(kp.order), _ptr(kp.count), F, mode, K, _ptr(hyp), seed,
                                                        float(thr), float(confidence), _ptr(r.H), _ptr(r.used_mask),
                                                        _ptr(r.inlier_mask), _ptr(r.status), _ptr(r.info), _ptr(sched), _ptr(retry),
                                                        _stream()), "egl_fit_homography_masked")
        return r

    # -- F4: keypoint propagation ----------------------------------------------------------
    def gray_pyramid(self, frames: torch.Tensor, max_level: int = 2, out: torch.Tensor | None = None) -> torch.Tensor:
        """frames (F, H, W, 3) uint8 BGR on the device -> (F, egl_pyramid_bytes) uint8 gray pyramids."""
        _require(frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_cuda,
                 "gray_pyramid: frames must be a CUDA uint8 tensor of shape (F, H, W, 3)")
        _require(frames.stride(3) == 1 and frames.stride(2) == 3, "gray_pyramid: pixels must be packed B,G,R bytes")
        F, H, W, _ = frames.shape
        nbytes = int(N.lib.egl_pyramid_bytes(H, W, max_level))
        _require(nbytes > 0, "gray_pyramid: bad frame size / max_level")
        if out is None:
            out = torch.empty((F, nbytes), dtype=torch.uint8, device=frames.device)
        with torch.cuda.device(frames.device):
            N.check(N.lib.egl_gray_pyramid(_ptr(frames), F, H, W, frames.stride(1), frames.stride(0) if F > 1 else H * frames.stride(1),
                                           max_level, _ptr(out), _stream()), "egl_gray_pyramid")
        return out

    def alloc_pyramid(self, n_frames: int, height: int, width: int, max_level: int = 2, device=None) -> torch.Tensor:
        nbytes = int(N.lib.egl_pyramid_bytes(height, width, max_level))
        _require(nbytes > 0, "alloc_pyramid: bad frame size / max_level")
        return torch.empty((n_frames, nbytes), dtype=torch.uint8, device=device or self.device)

    def gray_pyramid_step(self, frames: torch.Tensor, pyr: torch.Tensor, first: int, step: int, max_level: int = 2) -> None:
        """Pyramids of frames first, first + step, first + 2 step, ... of a contiguous (F, H, W, 3) clip into the same rows
        of pyr (F, egl_pyramid_bytes): one step of every chain at a time (egl_gray_pyramid_strided)."""
        _require(frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_cuda and frames.is_contiguous(),
                 "gray_pyramid_step: frames must be a contiguous CUDA uint8 tensor of shape (F, H, W, 3)")
        F, H, W, _ = frames.shape
        _require(pyr.dtype == torch.uint8 and pyr.is_contiguous() and pyr.shape[0] == F, "gray_pyramid_step: pyr must be (F, bytes) uint8")
        n = (F - first + step - 1) // step if first < F else 0
        if n <= 0:
            return
        with torch.cuda.device(frames.device):
            N.check(N.lib.egl_gray_pyramid_strided(_ptr(frames[first]), n, H, W, frames.stride(1), step * frames.stride(0), max_level,
                                                   _ptr(pyr[first]), step * pyr.stride(0), _stream()), "egl_gray_pyramid_strided")

    def track(self, pyr: torch.Tensor, height: int, width: int, prev: KeypointSet, prev0: int, next0: int, frame_step: int,
              max_level: int = 2, max_count: int = 10, eps: float = 0.03, out=None):
        """cv2.calcOpticalFlowPyrLK for n = prev.n_frames frame pairs (prev0 + p*step -> next0 + p*step).
        Returns (new_pts (n, 64, 2) float32, status (n, 64) uint8), indexed by channel."""
        n = prev.n_frames
        if out is not None:
            new_pts, status = out
            _require(new_pts.dtype == torch.float32 and new_pts.is_contiguous() and tuple(new_pts.shape) == (n, N.ORDER_STRIDE, 2)
                     and status.dtype == torch.uint8 and status.is_contiguous() and tuple(status.shape) == (n, N.ORDER_STRIDE),
                     "track: out must be ((n, 64, 2) float32, (n, 64) uint8)")
        else:
            new_pts = torch.empty((n, N.ORDER_STRIDE, 2), dtype=torch.float32, device=pyr.device)
            status = torch.empty((n, N.ORDER_STRIDE), dtype=torch.uint8, device=pyr.device)
        _require(pyr.dtype == torch.uint8 and pyr.is_contiguous() and pyr.dim() == 2
                 and pyr.shape[1] == int(N.lib.egl_pyramid_bytes(height, width, max_level)), "track: pyr does not match the frame size")
        last = max(prev0, next0) + (n - 1) * frame_step
        _require(0 <= min(prev0, next0) and last < pyr.shape[0], "track: frame index out of range")
        with torch.cuda.device(pyr.device):
            N.check(N.lib.egl_track_keypoints(_ptr(pyr), height, width, max_level, _ptr(prev.xy), _ptr(prev.order), _ptr(prev.count), n,
                                              prev0, next0, frame_step, max_count, float(eps), _ptr(new_pts), _ptr(status), _stream()),
                    "egl_track_keypoints")
        return new_pts, status

    def filter_flow(self, frames: torch.Tensor, hue0: int, frame_step: int, prev: KeypointSet, new_pts: torch.Tensor,
                    status: torch.Tensor, out: KeypointSet) -> KeypointSet:
        n = prev.n_frames
        F, H, W, _ = frames.shape
        _require(0 <= hue0 and hue0 + (n - 1) * frame_step < F, "filter_flow: frame index out of range")
        with torch.cuda.device(frames.device):
            N.check(N.lib.egl_filter_flow(_ptr(frames), H, W, frames.stride(1), frames.stride(0) if F > 1 else H * frames.stride(1), hue0,
                                          frame_step, _ptr(prev.xy), _ptr(prev.order), _ptr(prev.count), _ptr(new_pts), _ptr(status), n,
                                          _ptr(out.xy), _ptr(out.order), _ptr(out.count), _ptr(out.src), _stream()), "egl_filter_flow")
        return out

    def merge(self, a: KeypointSet, b: KeypointSet, apply: torch.Tensor | None = None) -> KeypointSet:
        """a = {**a, **b} per frame, in place."""
        with torch.cuda.device(a.xy.device):
            N.check(N.lib.egl_merge_keypoints(_ptr(a.xy), _ptr(a.order), _ptr(a.count), _ptr(a.src), _ptr(b.xy), _ptr(b.order),
                                              _ptr(b.count), _ptr(b.src), _ptr(apply), a.n_frames, _stream()), "egl_merge_keypoints")
        return a

    def calibrate(self, frames: torch.Tensor, frame0: int, frame_step: int, kp: KeypointSet, err: torch.Tensor) -> KeypointSet:
        F, H, W, _ = frames.shape
        n = kp.n_frames
        _require(0 <= frame0 and frame0 + (n - 1) * frame_step < F, "calibrate: frame index out of range")
        _require(err.dtype == torch.int32 and err.numel() == n, "calibrate: err must be (n,) int32")
        with torch.cuda.device(frames.device):
            N.check(N.lib.egl_calibrate_keypoints(_ptr(frames), H, W, frames.stride(1), frames.stride(0) if F > 1 else H * frames.stride(1),
                                                  frame0, frame_step, _ptr(kp.xy), _ptr(kp.order), _ptr(kp.count), _ptr(kp.src), _ptr(err),
                                                  n, _stream()), "egl_calibrate_keypoints")
        return kp

    def commit(self, kp: KeypointSet, fit: FitResult, sched: torch.Tensor | None, retry: torch.Tensor | None,
               fit_ok: torch.Tensor | None) -> None:
        with torch.cuda.device(kp.xy.device):
            N.check(N.lib.egl_commit_fit(_ptr(kp.order), _ptr(kp.count), _ptr(kp.src), _ptr(fit.status), _ptr(fit.inlier_mask),
                                         _ptr(sched), _ptr(retry), _ptr(fit_ok), kp.n_frames, _stream()), "egl_commit_fit")

    # -- cadence ----------------------------------------------------------------------------
    def select(self, status: torch.Tensor, interval: int = 1, carry_in: int = -1):
        """(h_index (F,) int32, attempted (F,) uint8) per the reference's homography cadence."""
        F = status.numel()
        h_index = torch.empty(F, dtype=torch.int32, device=status.device)
        attempted = torch.empty(F, dtype=torch.uint8, device=status.device)
        with torch.cuda.device(status.device):
            N.check(N.lib.egl_select_homography(_ptr(status), F, int(interval), int(carry_in), _ptr(h_index), _ptr(attempted),
                                                _stream()), "egl_select_homography")
        return h_index, attempted

    def select_chunk(self, status: torch.Tensor, interval: int, first_frame: int, carry: torch.Tensor,
                     h_index: torch.Tensor | None = None, attempted: torch.Tensor | None = None):
        """The cadence for one chunk of a clip: ``status`` belongs to clip frames first_frame.. (first_frame a multiple
        of interval); ``carry`` is a 1-element int32 device tensor holding the clip row of H valid before the chunk
        (-1 = none), updated in place.  h_index holds clip rows."""
        F = status.numel()
        _require(carry.dtype == torch.int32 and carry.numel() == 1 and carry.is_cuda, "select_chunk: carry must be a 1-element int32 CUDA tensor")
        if h_index is None:
            h_index = torch.empty(F, dtype=torch.int32, device=status.device)
        if attempted is None:
            attempted = torch.empty(F, dtype=torch.uint8, device=status.device)
        with torch.cuda.device(status.device):
            N.check(N.lib.egl_select_homography_chunk(_ptr(status), F, int(interval), int(first_frame), _ptr(carry), _ptr(carry),
                                                      _ptr(h_index), _ptr(attempted), _stream()), "egl_select_homography_chunk")
        return h_index, attempted

    # -- K4 ---------------------------------------------------------------------------------
    def alloc_projection(self, F: int, P: int) -> Projection:
        dev = self.device
        return Projection(torch.empty((F, P, 2), dtype=torch.float32, device=dev), torch.empty((F, P, 2), dtype=torch.int64, device=dev),
                          torch.empty((F, P), dtype=torch.uint8, device=dev), torch.empty((F, 4), dtype=torch.float64, device=dev))

    def project(self, H: torch.Tensor, foot: torch.Tensor, count: torch.Tensor, img_w: int, img_h: int,
                h_index: torch.Tensor | None = None, out: Projection | None = None) -> Projection:
        """foot (F, P, 2) float32, count (F,) int32, H (*, 9) float64; h_index (F,) int32 or None."""
        _require(foot.dtype == torch.float32 and count.dtype == torch.int32 and H.dtype == torch.float64,
                 "project: dtypes must be foot float32, count int32, H float64")
        _require(foot.is_contiguous() and H.is_contiguous() and foot.dim() == 3, "project: foot (F, P, 2) and H must be contiguous")
        F, P, _ = foot.shape
        pr = out if out is not None else self.alloc_projection(F, P)
        if h_index is not None:
            _require(h_index.dtype == torch.int32 and h_index.numel() == F, "project: h_index must be int32 with one entry per frame")
        with torch.cuda.device(foot.device):
            N.check(N.lib.egl_project_points(_ptr(H), _ptr(h_index), _ptr(foot), _ptr(count), F, P, img_w, img_h, _ptr(pr.coords),
                                             _ptr(pr.coords_i), _ptr(pr.in_bounds), _ptr(pr.bounds), _stream()),
                    "egl_project_points")
        return pr
