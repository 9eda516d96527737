"""Decode front end with the reference's ``KeypointModel.get_keypoints`` signature
(eagle/models/keypoint_hrnet.py:575-595): tensor in, ``list[list[(channel, x_n, y_n, score)]]`` out.

The reference copies each of the 57 heatmaps of every frame to the host and calls ``np.argmax`` on
it; here one CUDA kernel reduces all maps and only (index, score) pairs come back."""
from __future__ import annotations

from typing import Callable

import torch

from .engine import GeometryEngine


class KeypointDecoder:
    def __init__(self, model: Callable | None = None, device="cuda:0"):
        """model: optional callable x (N,3,H,W) -> heatmaps (N,57,h,w) (the HRNet forward incl. sigmoid);
        without it ``get_keypoints`` takes the heatmaps themselves."""
        self.model = model
        self.engine = GeometryEngine(device)

    @torch.no_grad()
    def get_keypoints(self, x: torch.Tensor):
        heatmaps = self.model(x) if self.model is not None else x
        heatmaps = heatmaps.to(self.engine.device, torch.float32).contiguous()
        N, C, H, W = heatmaps.shape
        kp = self.engine.decode(heatmaps, 1, 1)  # image size is irrelevant for the normalised output
        flat = kp.flat.cpu().numpy()
        score = kp.score.cpu().numpy()
        out = []
        for n in range(N):
            coords = []
            for i in range(C):
                y, xx = divmod(int(flat[n, i]), W)
                s = float(score[n, i])
                if s > 0.01:  # keypoint_hrnet.py:592
                    coords.append((i, xx / max(1, W - 1), y / max(1, H - 1), s))
            out.append(coords)
        return out
