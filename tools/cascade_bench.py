#!/usr/bin/env python
"""Cost of the RHO / LMEDS legs on the GPU: egl_fit_homography on (a) the 360 golden cascade sets, (b) a 2250-frame clip in
which k frames have no RANSAC model (points on one pitch line plus strays), k = 0, 1, 16, 256.  CUDA-event times."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.engine import GeometryEngine, KeypointSet  # noqa: E402


def kp_from_lists(channels, img_pts, counts):
    T = len(counts)
    xy = np.zeros((T, 57, 2), np.int32); order = np.full((T, 64), 255, np.uint8); count = np.zeros((T, 2), np.int32)
    for i in range(T):
        n = int(counts[i]); ch = channels[i, :n]
        xy[i, ch] = img_pts[i, :n].astype(np.int32); order[i, :n] = ch; count[i] = n
    return xy, order, count


def to_kp(xy, order, count):
    T = len(count)
    return KeypointSet(torch.zeros((T, 57), dtype=torch.int32).cuda(), torch.zeros((T, 57)).cuda(), torch.from_numpy(xy).cuda(),
                       torch.from_numpy(order).cuda(), torch.from_numpy(count).cuda())


def timed(eng, kp, reps=20):
    fit = eng.alloc_fit(kp.n_frames)
    for _ in range(3):
        eng.fit(kp, out=fit)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.fit(kp, out=fit)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, fit


eng = GeometryEngine("cuda:0")
g = np.load(os.path.join(ROOT, "tests", "golden", "cascade_cv2.npz"))
gx, go, gc = kp_from_lists(g["channels"], g["img_pts"], g["n"])
out = {}
ms, fit = timed(eng, to_kp(gx, go, gc))
info = fit.info.cpu().numpy()
out["golden_360_sets_ms"] = ms
out["golden_legs"] = {"rho": int((info[:, 2] == -2).sum()), "lmeds": int((info[:, 2] == -3).sum()), "none": int((fit.status.cpu().numpy() == 2).sum())}
# a clip of ordinary frames with k hard ones mixed in
clip = synthetic.make_clip(2250, 1920, 1080, seed=3, ghost_prob=0.05)
kp0 = eng.synthesize(eng.decode(torch.from_numpy(clip["heatmaps"]).cuda(), 1920, 1080))
xy = kp0.xy.cpu().numpy(); order = kp0.order.cpu().numpy(); count = kp0.count.cpu().numpy()
hard = np.where(g["leg"] != 0)[0]
for k in (0, 1, 16, 256):
    x2, o2, c2 = xy.copy(), order.copy(), count.copy()
    for j in range(k):
        src = hard[j % len(hard)]; dst = (j * 2250) // max(k, 1)
        x2[dst], o2[dst], c2[dst] = gx[src], go[src], gc[src]
    ms, fit = timed(eng, to_kp(x2, o2, c2), reps=10)
    out[f"clip_2250_with_{k}_hard_frames_ms"] = ms
print(json.dumps(out, indent=1))
