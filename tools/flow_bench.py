"""Time the sparse-keypoint-cadence path (eagle_b200.propagation) on one GPU: main.py's cadence
(fps 25, num_homography=1, num_keypoint_detection=3 -> network every 8th frame, Lucas-Kanade in between).

Synthetic clip: a few rendered 8-frame camera moves (540p renders upsampled to 1080p on the device), tiled
to F frames with per-frame noise so that no two frames are equal.  Frames and head heatmaps are resident
in HBM when the timed region starts.  Prints one JSON line with the per-stage split.

    python tools/flow_bench.py [--frames 2250] [--steps 5] [--paths 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_clip(torch, dev, F, k, paths, seed=0):
    from eagle_b200 import synthetic
    return synthetic.tiled_flow_clip_device(F, k, dev, paths, seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2250)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--paths", type=int, default=3)
    ap.add_argument("--fps", type=int, default=25)
    ap.add_argument("--calibration", action="store_true")
    args = ap.parse_args()
    import torch
    from eagle_b200.engine import GeometryEngine
    from eagle_b200.propagation import PropagatedPath
    dev = torch.device("cuda:0")
    e = GeometryEngine(dev)
    k = max(1, int(args.fps / 3)); h = max(1, int(args.fps / 1))
    t0 = time.time()
    frames, heads = build_clip(torch, dev, args.frames, k, args.paths)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    prop = PropagatedPath(e)
    foot = torch.rand((args.frames, 27, 2), device=dev) * torch.tensor([1920.0, 1080.0], device=dev)
    cnt = torch.full((args.frames,), 25, dtype=torch.int32, device=dev)

    def once():
        out = prop.run(frames, heads, None, k, h, args.calibration)
        e.project(out["H"], foot, cnt, 1920, 1080, h_index=out["h_index"])
        return out

    for _ in range(2):
        out = once()
    torch.cuda.synchronize()
    # stage split with events (one extra run)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    a, b = ev(), ev()
    a.record(); e.gray_pyramid(frames, 2); b.record(); torch.cuda.synchronize()
    pyr_ms = a.elapsed_time(b)
    t0 = time.perf_counter()
    s0, s1 = ev(), ev()
    s0.record()
    for _ in range(args.steps):
        out = once()
    s1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps
    dev_ms = s0.elapsed_time(s1) / args.steps
    kpf = out["count"][:, 0].float().mean().item()
    line = {"workload": f"{args.frames} frames 1080p, keypoint interval {k}, homography interval {h}, calibration {args.calibration}",
            "frames_per_s_wall": args.frames / wall, "ms_per_clip_wall": wall * 1e3, "ms_per_clip_device": dev_ms,
            "gray_pyramid_ms": pyr_ms, "gray_pyramid_GBps": args.frames * (1080 * 1920 * 3 + 2721600) / (pyr_ms * 1e-3) / 1e9,
            "mean_keypoints_per_frame": kpf, "stats": prop.stats, "frames_with_H": int((out["h_index"] >= 0).sum().item()),
            "clip_build_s": build_s}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
