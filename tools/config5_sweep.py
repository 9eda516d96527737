"""BASELINE.json configs[4]: 4K (3840x2160) frames -- preprocessing + heatmap decode bandwidth sweep over
batch sizes 1..256 (CUDA events, L2 flushed between timed launches by cycling distinct buffers)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200.engine import GeometryEngine  # noqa: E402

eng = GeometryEngine("cuda:0")
PEAK = 6545.3


def timeit(fn, n_bufs, iters=12, warm=3):
    for i in range(warm):
        fn(i % n_bufs)
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i % n_bufs)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


print("F, K1 4K ms, K1 GB/s (touched rows), K1 frac, K2 ms, K2 GB/s, K2 frac")
for F in [1, 2, 4, 8, 16, 32, 64, 128, 256]:
    # enough distinct buffers that consecutive launches never hit L2 (126 MB)
    nb = max(2, min(16, int(400e6 // (F * 2160 * 3840 * 3)) + 2))
    frames = [torch.randint(0, 256, (F, 2160, 3840, 3), dtype=torch.uint8, device="cuda") for _ in range(nb)]
    out = torch.empty((F, 3, 540, 960), device="cuda")
    t1 = timeit(lambda i: eng.preprocess(frames[i], out=out), nb)
    alg1 = F * (2160 // 2 * 3840 * 3 + 3 * 540 * 960 * 4)
    del frames
    nbh = max(2, min(16, int(400e6 // (F * 57 * 135 * 240 * 4)) + 2))
    hms = [torch.rand((F, 57, 135, 240), device="cuda") for _ in range(nbh)]
    kp = eng.alloc_keypoints(F)
    t2 = timeit(lambda i: eng.decode(hms[i], 3840, 2160, out=kp), nbh)
    alg2 = F * 57 * 135 * 240 * 4
    del hms
    print(f"{F}, {t1:.4f}, {alg1 / t1 / 1e6:.0f}, {alg1 / t1 / 1e6 / PEAK:.2f}, {t2:.4f}, {alg2 / t2 / 1e6:.0f}, {alg2 / t2 / 1e6 / PEAK:.2f}")
