#!/usr/bin/env python
"""Cost of the <= 9-inlier refit path (lane 0 runs the scalar refit with OpenCV's eigenvalue-thresholded LM solves): a
2250-frame 1080p fit in which k frames keep only 7 on-plane landmarks (all inliers), k = 0, 1, 64, 2250."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import synthetic
from eagle_b200.engine import GeometryEngine, KeypointSet
eng = GeometryEngine("cuda:0")
clip = synthetic.make_clip(2250, 1920, 1080, seed=3, ghost_prob=0.0)
kp0 = eng.synthesize(eng.decode(torch.from_numpy(clip["heatmaps"]).cuda(), 1920, 1080))
xy = kp0.xy.cpu().numpy(); order = kp0.order.cpu().numpy(); count = kp0.count.cpu().numpy()
OFF = (0, 1, 24, 25)
out = {}
for k in (0, 1, 64, 2250):
    o2, c2 = order.copy(), count.copy()
    for j in range(k):
        f = (j * 2250) // max(k, 1)
        on = [int(c) for c in order[f, :count[f, 0]] if int(c) not in OFF]
        keep = on[:: max(1, len(on) // 7)][:7]
        o2[f] = 255; o2[f, :len(keep)] = keep; c2[f] = len(keep)
    kp = KeypointSet(kp0.flat, kp0.score, torch.from_numpy(xy).cuda(), torch.from_numpy(o2).cuda(), torch.from_numpy(c2).cuda())
    fit = eng.alloc_fit(2250)
    for _ in range(3): eng.fit(kp, out=fit)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): eng.fit(kp, out=fit)
    b.record(); torch.cuda.synchronize()
    info = fit.info.cpu().numpy()
    out[f"fit_ms_with_{k}_seven_point_frames"] = a.elapsed_time(b) / 10
    out[f"frames_with_le9_inliers_{k}"] = int((info[:, 1] <= 9).sum())
print(json.dumps(out, indent=1))
