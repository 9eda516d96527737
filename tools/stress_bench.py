"""RANSAC stress (BASELINE.json configs[3]): K=4096 hypotheses/frame, 53 landmarks, 40% outliers, 50k frames."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import _native as N  # noqa: E402
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.engine import GeometryEngine, KeypointSet  # noqa: E402

F = int(os.environ.get("F", 50000)); K = int(os.environ.get("K", 4096))
eng = GeometryEngine("cuda:0")
xy, valid, flags, cams = synthetic.stress_point_sets(256, 1920, 1080, seed=1)
reps = (F + 255) // 256
xy = np.tile(xy, (reps, 1, 1))[:F]; flags = np.tile(flags, (reps, 1))[:F]
on = [i for i in range(57) if i not in (0, 1, 24, 25)]
order = np.full((F, 64), 255, np.uint8); order[:, :53] = on
kp = KeypointSet(torch.zeros((F, 57), dtype=torch.int32).cuda(), torch.zeros((F, 57)).cuda(), torch.from_numpy(xy).cuda(),
                 torch.from_numpy(order).cuda(), torch.from_numpy(np.full((F, 2), 53, np.int32)).cuda())
fit = eng.alloc_fit(F)
for _ in range(2):
    eng.fit(kp, mode=N.FIT_FIXED_K, K=K, seed=1, out=fit)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
iters = 3
s.record()
for _ in range(iters):
    eng.fit(kp, mode=N.FIT_FIXED_K, K=K, seed=1, out=fit)
e.record(); torch.cuda.synchronize()
t = s.elapsed_time(e) / iters
flop = F * K * (627 + 21 * 53)
inl = fit.inlier_mask.cpu().numpy()
want = np.array([sum(1 << c for c in on if not flags[f, c]) for f in range(F)], dtype=np.int64)
print(f"fixed-K fit (hypotheses + refit) F={F} K={K}: {t:.3f} ms  {F / t * 1e3:.0f} frames/s  {flop / t / 1e9:.2f} TFLOP/s algorithmic; "
      f"frames whose inlier set == planted inliers: {(inl == want).mean() * 100:.2f}%  status ok: {(fit.status == 0).float().mean().item() * 100:.2f}%")
