"""Seeded synthetic workloads for the geometry path (frames, heatmaps, boxes, point sets).

There are no datasets or checkpoints offline, so every test and benchmark runs on synthetic data
of the reference's shapes: a broadcast-like camera homography per frame,
Gaussian landmark peaks rendered into (57, 135, 240) float32 heatmaps at the positions
``KeypointModel`` would emit them (keypoint_hrnet.py:590-591: x/(W-1), y/(H-1) normalisation), and
detector-shaped boxes whose foot points fall on the pitch.  Generation is numpy on the host (tests,
CPU baseline, the benchmark's pool of landmark layouts); bench.py adds per-frame noise on the device.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .pitch import NUM_LANDMARKS, OFF_PLANE, PITCH_LENGTH_M, PITCH_WIDTH_M, WORLD_XYZ

HM_H, HM_W = 135, 240  # HRNet-W48 branch-0 resolution for a 540x960 input (keypoint_hrnet.py:481)

# image(1280x720) -> pitch homography of a typical main-camera view
_BASE_IMG_TO_PITCH_720P = np.array([[0.09, -0.02, -5.0], [0.004, 0.16, -20.0], [1e-5, 6e-4, 1.0]])


def sample_cameras(n: int, width: int, height: int, rng: np.random.Generator) -> np.ndarray:
    """(n, 3, 3) float64 pitch->image homographies: pans/zooms/perturbations of the base view."""
    base = np.linalg.inv(_BASE_IMG_TO_PITCH_720P)  # pitch -> 720p image
    S = np.diag([width / 1280.0, height / 720.0, 1.0])
    out = np.empty((n, 3, 3))
    for i in range(n):
        pan = np.eye(3)
        pan[0, 2] = rng.uniform(-420.0, 420.0)  # px at 720p
        pan[1, 2] = rng.uniform(-40.0, 60.0)
        z = rng.uniform(0.8, 1.5)
        zoom = np.array([[z, 0, 640.0 * (1 - z)], [0, z, 360.0 * (1 - z)], [0, 0, 1.0]])
        pert = np.eye(3) + rng.normal(0.0, 1.0, (3, 3)) * np.array([[2e-2, 2e-2, 0], [2e-2, 2e-2, 0], [2e-6, 2e-6, 0]])
        H = S @ zoom @ pan @ base @ pert
        out[i] = H / H[2, 2]
    return out


def project_points(H: np.ndarray, pts: np.ndarray) -> np.ndarray:
    p = np.c_[pts, np.ones(len(pts))] @ H.T
    return p[:, :2] / p[:, 2:3]


def landmark_pixels(cam: np.ndarray, width: int, height: int, margin: float = 4.0):
    """Image positions of the 57 landmarks under ``cam`` and a visibility mask (inside the frame).

    Off-plane landmarks (cross-bar ends) are drawn above their goal-line foot; they are never
    used for the fit (pitch.OFF_PLANE) but do appear in the heatmaps like any other channel.
    """
    px = project_points(cam, WORLD_XYZ[:, :2])
    for ch in OFF_PLANE:
        px[ch, 1] -= 0.06 * height
    vis = (px[:, 0] >= margin) & (px[:, 0] <= width - 1 - margin) & (px[:, 1] >= margin) & (px[:, 1] <= height - 1 - margin)
    p = np.c_[WORLD_XYZ[:, :2], np.ones(NUM_LANDMARKS)] @ cam.T
    vis &= p[:, 2] > 0
    return px, vis


def render_heatmaps(px: np.ndarray, vis: np.ndarray, width: int, height: int, rng: np.random.Generator,
                    peak: float = 0.9, sigma: float = 1.5, background: float = 0.05, jitter: float = 0.3) -> np.ndarray:
    """(57, 135, 240) float32: U(0, background) noise plus one Gaussian bump per visible landmark."""
    hm = rng.uniform(0.0, background, (NUM_LANDMARKS, HM_H, HM_W)).astype(np.float32)
    ys, xs = np.mgrid[0:HM_H, 0:HM_W]
    for ch in np.nonzero(vis)[0]:
        cx = px[ch, 0] / width * (HM_W - 1) + rng.normal(0.0, jitter)
        cy = px[ch, 1] / height * (HM_H - 1) + rng.normal(0.0, jitter)
        g = peak * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2.0 * sigma * sigma))
        hm[ch] += g.astype(np.float32)
    np.clip(hm, 0.0, 1.0, out=hm)
    return hm


def blank_heatmaps(heatmaps: np.ndarray, frames: Sequence[int]) -> None:
    """Replace the given frames' heatmaps, in place, by four sharp peaks on the cross-bar channels (pitch.OFF_PLANE).
    Such a frame still decodes four landmarks (so there is no optical-flow rescue, coordinate_model.py:287) but has no
    on-plane correspondence at all: its homography fit fails (:350-352) and the reference retries on the next frame."""
    for f in frames:
        heatmaps[f] = 0.0
        for j, c in enumerate(OFF_PLANE):
            heatmaps[f, c, 20 + 10 * j, 30 + 15 * j] = 0.9


def sample_objects(cam: np.ndarray, width: int, height: int, rng: np.random.Generator,
                   n_players: int = 20, n_goalkeepers: int = 2, ball: bool = True) -> dict:
    """A ``detect_objects``-shaped dict (coordinate_model.py:557-628): clipped int boxes, confidences,
    ``Bottom_center = [int((x1+x2)/2), y2]``.  Most foot points fall on the visible pitch; a few
    are placed off it so that the out-of-bounds branch (:385-386) is exercised."""
    out = {"Player": {}, "Goalkeeper": {}}
    inv = np.linalg.inv(cam)
    s = height / 720.0
    next_id = 1

    def one_box(w_px, h_px):
        for _ in range(50):
            foot = np.array([rng.uniform(0.05, 0.95) * width, rng.uniform(0.30, 0.97) * height])
            wp = project_points(inv, foot[None])[0]
            if rng.uniform() < 0.08 or (0 <= wp[0] <= PITCH_LENGTH_M and 0 <= wp[1] <= PITCH_WIDTH_M):
                break
        x1 = int(np.clip(foot[0] - w_px / 2, 0, width - 1)); x2 = int(np.clip(foot[0] + w_px / 2, 0, width - 1))
        y2 = int(np.clip(foot[1], 0, height - 1)); y1 = int(np.clip(foot[1] - h_px, 0, height - 1))
        return [x1, y1, x2, y2]

    for cls, count in (("Player", n_players), ("Goalkeeper", n_goalkeepers)):
        for _ in range(count):
            box = one_box(rng.uniform(30, 50) * s, rng.uniform(70, 110) * s)
            out[cls][next_id] = {"BBox": box, "Confidence": float(rng.uniform(0.4, 0.95)),
                                 "Bottom_center": [int((box[0] + box[2]) / 2), box[3]]}
            next_id += 1
    if ball:
        b = one_box(10 * s, 10 * s)
        box = np.array(b)  # the reference keeps the ball box as an int ndarray (:622,627)
        out["Ball"] = {0: {"BBox": box, "Confidence": float(rng.uniform(0.4, 0.9)),
                           "Bottom_center": [int((box[0] + box[2]) / 2), box[3]]}}
    return out


def make_clip(n_frames: int, width: int, height: int, seed: int = 0, with_frames: bool = False,
              hide_prob: float = 0.1, ghost_prob: float = 0.0):
    """A synthetic clip: dict(cameras, heatmaps (F,57,135,240) f32, objects list, frames or None).

    hide_prob drops a visible landmark's bump (missed detection); ghost_prob adds a bump for a
    landmark at a wrong place (outlier correspondence).
    """
    rng = np.random.default_rng(seed)
    cams = sample_cameras(n_frames, width, height, rng)
    heat = np.empty((n_frames, NUM_LANDMARKS, HM_H, HM_W), np.float32)
    objs = []
    for i in range(n_frames):
        px, vis = landmark_pixels(cams[i], width, height)
        vis = vis & (rng.uniform(size=NUM_LANDMARKS) >= hide_prob)
        if ghost_prob > 0:
            ghosts = rng.uniform(size=NUM_LANDMARKS) < ghost_prob
            px = px.copy()
            px[ghosts] = rng.uniform([8, 8], [width - 9, height - 9], (int(ghosts.sum()), 2))
            vis = vis | ghosts
        heat[i] = render_heatmaps(px, vis, width, height, rng)
        objs.append(sample_objects(cams[i], width, height, rng))
    frames = None
    if with_frames:
        frames = rng.integers(0, 256, (n_frames, height, width, 3), dtype=np.uint8)
    return {"cameras": cams, "heatmaps": heat, "objects": objs, "frames": frames, "width": width, "height": height}


from .boxes import objects_to_arrays  # noqa: E402,F401  (re-exported: the packing belongs to the product API)


def stress_point_sets(n_frames: int, width: int, height: int, seed: int = 0, outlier_frac: float = 0.4,
                      noise_px: float = 0.5):
    """RANSAC stress inputs (BASELINE.json configs[3]): all 53 on-plane landmarks present, a fraction
    replaced by gross outliers (>= 15 m off in pitch space), inliers with sub-pixel noise, rounded to
    integer pixels.  Returns (kp_xy (F,57,2) int32, valid mask (F,) uint64, outlier flags (F,57) bool,
    cameras)."""
    rng = np.random.default_rng(seed)
    on = np.array([i for i in range(NUM_LANDMARKS) if i not in OFF_PLANE])
    n_out = int(round(outlier_frac * len(on)))
    kp = np.zeros((n_frames, NUM_LANDMARKS, 2), np.int32)
    flags = np.zeros((n_frames, NUM_LANDMARKS), bool)
    cams = np.empty((n_frames, 3, 3))
    base = np.linalg.inv(_BASE_IMG_TO_PITCH_720P)
    S = np.diag([width / 1280.0, height / 720.0, 1.0])
    for f in range(n_frames):
        # a wide view that keeps every landmark in front of the camera
        pert = np.eye(3) + rng.normal(0.0, 1.0, (3, 3)) * np.array([[1e-2, 1e-2, 0], [1e-2, 1e-2, 0], [1e-6, 1e-6, 0]])
        cam = S @ base @ pert
        cam /= cam[2, 2]
        cams[f] = cam
        px = project_points(cam, WORLD_XYZ[:, :2]) + rng.normal(0.0, noise_px, (NUM_LANDMARKS, 2))
        inv = np.linalg.inv(cam)
        out_idx = rng.choice(on, n_out, replace=False)
        for ch in out_idx:
            for _ in range(100):
                cand = rng.uniform([0, 0], [width, height])
                wp = project_points(inv, cand[None])[0]
                if np.hypot(*(wp - WORLD_XYZ[ch, :2])) >= 15.0:
                    break
            px[ch] = cand
        flags[f, out_idx] = True
        kp[f] = np.rint(px).astype(np.int32)
    valid = np.full(n_frames, sum(1 << int(i) for i in on), np.uint64)
    return kp, valid, flags, cams


# --------------------------------------------------------------------------------------------------
# rendered clips: smooth camera motion + pitch markings, for the keypoint-propagation (optical flow) path
# --------------------------------------------------------------------------------------------------
def camera_path(n: int, width: int, height: int, rng: np.random.Generator, pan_px: float = 3.0, zoom_rate: float = 0.002) -> np.ndarray:
    """(n, 3, 3) pitch->image homographies of one camera panning / zooming smoothly (a few px per frame)."""
    base = np.linalg.inv(_BASE_IMG_TO_PITCH_720P)
    S = np.diag([width / 1280.0, height / 720.0, 1.0])
    x0 = rng.uniform(-150.0, 150.0); y0 = rng.uniform(-20.0, 30.0); z0 = rng.uniform(0.95, 1.2)
    vx = rng.uniform(-pan_px, pan_px); vy = rng.uniform(-0.3, 0.3) * pan_px; vz = rng.uniform(-zoom_rate, zoom_rate)
    out = np.empty((n, 3, 3))
    for i in range(n):
        # slow sinusoidal variation of the pan speed so that the motion is not exactly uniform
        px = x0 + vx * i + 4.0 * np.sin(0.21 * i); py = y0 + vy * i + 1.5 * np.sin(0.13 * i + 1.0)
        z = z0 + vz * i
        pan = np.eye(3); pan[0, 2] = px; pan[1, 2] = py
        zoom = np.array([[z, 0, 640.0 * (1 - z)], [0, z, 360.0 * (1 - z)], [0, 0, 1.0]])
        H = S @ zoom @ pan @ base
        out[i] = H / H[2, 2]
    return out


def _marking_distance(wx: np.ndarray, wy: np.ndarray) -> np.ndarray:
    """Distance in metres from pitch points to the nearest marking of a standard 105 x 68 pitch."""
    L, Wd = float(PITCH_LENGTH_M), float(PITCH_WIDTH_M)
    d = np.full(wx.shape, 1e9)

    def vseg(x, y0, y1):
        nonlocal d
        dy = np.maximum(np.maximum(y0 - wy, wy - y1), 0.0)
        d = np.minimum(d, np.hypot(wx - x, dy))

    def hseg(y, x0, x1):
        nonlocal d
        dx = np.maximum(np.maximum(x0 - wx, wx - x1), 0.0)
        d = np.minimum(d, np.hypot(dx, wy - y))

    vseg(0.0, 0.0, Wd); vseg(L, 0.0, Wd); vseg(L / 2, 0.0, Wd); hseg(0.0, 0.0, L); hseg(Wd, 0.0, L)
    for x_goal, sgn in ((0.0, 1.0), (L, -1.0)):
        for depth, half in ((16.5, 20.16), (5.5, 9.16)):
            xe = x_goal + sgn * depth
            vseg(xe, Wd / 2 - half, Wd / 2 + half)
            hseg(Wd / 2 - half, min(x_goal, xe), max(x_goal, xe)); hseg(Wd / 2 + half, min(x_goal, xe), max(x_goal, xe))
        # penalty arc (outside the box only) and spot
        sx = x_goal + sgn * 11.0
        r = np.hypot(wx - sx, wy - Wd / 2)
        outside = (sgn * (wx - x_goal)) > 16.5
        d = np.minimum(d, np.where(outside, np.abs(r - 9.15), 1e9))
        d = np.minimum(d, np.maximum(r - 0.15, 0.0))
    d = np.minimum(d, np.abs(np.hypot(wx - L / 2, wy - Wd / 2) - 9.15))
    return d


def render_pitch_frame(cam: np.ndarray, width: int, height: int, rng: np.random.Generator | None = None, noise: float = 1.5) -> np.ndarray:
    """(height, width, 3) uint8 BGR picture of the pitch seen through ``cam``: mown grass with a texture
    that is fixed in pitch coordinates (so it moves with the camera), white markings, grey surroundings."""
    inv = np.linalg.inv(cam)
    ys, xs = np.mgrid[0:height, 0:width].astype(np.float64)
    den = inv[2, 0] * xs + inv[2, 1] * ys + inv[2, 2]
    wx = (inv[0, 0] * xs + inv[0, 1] * ys + inv[0, 2]) / den
    wy = (inv[1, 0] * xs + inv[1, 1] * ys + inv[1, 2]) / den
    tex = 7.0 * np.sin(1.31 * wx + 0.73 * wy) + 5.0 * np.sin(2.9 * wy - 1.1 * wx + 0.4) + 4.0 * np.sin(0.47 * wx * wy * 0.05 + 1.7)
    stripe = np.where((np.floor(wx / 5.25).astype(np.int64) & 1) == 0, 10.0, -10.0)
    inside = (wx >= -2.0) & (wx <= PITCH_LENGTH_M + 2.0) & (wy >= -2.0) & (wy <= PITCH_WIDTH_M + 2.0) & (den > 0)
    img = np.empty((height, width, 3), np.float64)
    img[..., 0] = np.where(inside, 45.0 + 0.6 * tex, 95.0 + tex)                 # B
    img[..., 1] = np.where(inside, 125.0 + stripe + tex, 100.0 + 1.2 * tex)      # G
    img[..., 2] = np.where(inside, 50.0 + 0.8 * tex, 105.0 + 0.9 * tex)          # R
    d = _marking_distance(wx, wy)
    line = np.clip(1.5 - d / 0.12, 0.0, 1.0) * inside  # ~0.3 m wide, soft edge
    img += (235.0 - img) * line[..., None]
    if rng is not None and noise > 0:
        img += rng.normal(0.0, noise, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_flow_clip(n_frames: int, width: int, height: int, seed: int = 0, hide_prob: float = 0.1, pan_px: float = 3.0):
    """Like make_clip, but with one smoothly moving camera and rendered frames that show the markings
    the heatmaps point at -- what the Lucas-Kanade propagation needs."""
    rng = np.random.default_rng(seed)
    cams = camera_path(n_frames, width, height, rng, pan_px=pan_px)
    heat = np.empty((n_frames, NUM_LANDMARKS, HM_H, HM_W), np.float32)
    frames = np.empty((n_frames, height, width, 3), np.uint8)
    objs = []
    for i in range(n_frames):
        px, vis = landmark_pixels(cams[i], width, height)
        vis = vis & (rng.uniform(size=NUM_LANDMARKS) >= hide_prob)
        heat[i] = render_heatmaps(px, vis, width, height, rng)
        objs.append(sample_objects(cams[i], width, height, rng))
        frames[i] = render_pitch_frame(cams[i], width, height, rng)
    return {"cameras": cams, "heatmaps": heat, "objects": objs, "frames": frames, "width": width, "height": height}


def tiled_flow_clip_device(n_frames: int, keypoint_interval: int, device, paths: int = 3, seed: int = 0, out_frames=None):
    """A long 1080p clip for timing the keypoint-propagation path, built on the device: ``paths`` rendered
    camera moves of ``keypoint_interval`` frames each (540p renders, upsampled 2x), tiled chain after chain
    with +-2 levels of per-frame noise so that no two frames are equal.  Returns (frames (F,1080,1920,3) uint8,
    head heatmaps (ceil(F/k),57,135,240) float32), both CUDA tensors."""
    import torch
    k = keypoint_interval
    pool_f, pool_h = [], []
    for p in range(paths):
        c = make_flow_clip(k, 960, 540, seed=seed + p, pan_px=1.5)
        fr = torch.from_numpy(c["frames"]).to(device).permute(0, 3, 1, 2).float()
        fr = torch.nn.functional.interpolate(fr, scale_factor=2, mode="bilinear", align_corners=False)
        pool_f.append(fr.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous())
        pool_h.append(torch.from_numpy(c["heatmaps"][0:1]).to(device))
    nc = (n_frames + k - 1) // k
    frames = out_frames if out_frames is not None else torch.empty((n_frames, 1080, 1920, 3), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device); g.manual_seed(seed)
    for c in range(nc):
        n = min(k, n_frames - c * k)
        src = pool_f[c % paths][:n]
        noise = torch.randint(-2, 3, src.shape, device=device, generator=g, dtype=torch.int16)
        frames[c * k:c * k + n] = (src.to(torch.int16) + noise).clamp(0, 255).to(torch.uint8)
    heads = torch.cat([pool_h[c % paths] for c in range(nc)])
    heads = (heads + torch.rand(heads.shape, device=device, generator=g) * 0.01).clamp(0, 1).contiguous()
    return frames, heads
