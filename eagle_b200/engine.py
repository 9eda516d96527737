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

    @property
    def n_frames(self) -> int:
        return self.flat.shape[0]


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


class GeometryEngine:
    """Thin, stateless-per-call wrapper; ``device`` must be a CUDA device."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise N.NativeError("GeometryEngine needs a CUDA device; there is no CPU path")

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
            seed: int = 0, thr: float = 5.0, confidence: float = 0.995, out: FitResult | None = None) -> FitResult:
        F = kp.n_frames
        r = out if out is not None else self.alloc_fit(F)
        if hyp is not None:
            _require(hyp.dtype == torch.uint8 and hyp.is_cuda and hyp.is_contiguous() and tuple(hyp.shape) == (F, K, 4),
                     "fit: hyp must be a contiguous CUDA uint8 tensor of shape (F, K, 4)")
        with torch.cuda.device(kp.xy.device):
            N.check(N.lib.egl_fit_homography(_ptr(kp.xy), _ptr(kp.order), _ptr(kp.count), F, mode, K, _ptr(hyp), seed, float(thr),
                                             float(confidence), _ptr(r.H), _ptr(r.used_mask), _ptr(r.inlier_mask), _ptr(r.status),
                                             _ptr(r.info), _stream()), "egl_fit_homography")
        return r

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
