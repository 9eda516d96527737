#!/usr/bin/env python
"""egl_upload_frames (cache-resident pinned ring) against the two alternatives on one GPU: a thread pool copying whole
frames into a large page-locked buffer followed by one H2D, and the same with both legs overlapped chunk by chunk."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200.engine import GeometryEngine
H, W, NF = 1080, 1920, 256
eng = GeometryEngine("cuda:0")
frames = [np.random.randint(0, 255, (H, W, 3), dtype=np.uint8) for _ in range(NF)]
dev = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
nb = frames[0].nbytes
for nt in (2, 4, 6, 8, 12, 16):
    eng.upload_frames(frames[:32], dev, threads=nt)
    t0 = time.perf_counter()
    for rep in range(3):
        eng.upload_frames(frames, dev, threads=nt)
    dt = (time.perf_counter() - t0) / 3
    print(f"egl_upload_frames {nt:2d} threads: {1e6 * dt / NF:.0f} us/frame ({NF * nb / dt / 1e9:.1f} GB/s)")
ok = all(np.array_equal(dev[i].cpu().numpy(), frames[i]) for i in (0, 1, 17, NF - 1))
print("bytes identical:", ok)
# reference points: big pinned staging
stage = torch.empty((64, H, W, 3), dtype=torch.uint8, pin_memory=True)
sn = stage.numpy()
for nt in (8, 16):
    with ThreadPoolExecutor(nt) as pool:
        t0 = time.perf_counter()
        for c in range(0, NF, 64):
            list(pool.map(lambda j: np.copyto(sn[j], frames[c + j]), range(64)))
            dev[c:c + 64].copy_(stage, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"pool of {nt} -> 64-frame pinned buffer -> H2D, serial: {1e6 * dt / NF:.0f} us/frame ({NF * nb / dt / 1e9:.1f} GB/s)")
