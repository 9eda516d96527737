"""How fast can frames that live in ordinary (pageable) numpy arrays be page-locked in place?  Decides whether the
drop-in API stages frames through its own pinned buffers (a host memcpy per frame) or registers the caller's memory."""
import sys
import time

import numpy as np
import torch

rt = torch.cuda.cudart()
torch.cuda.init()
H, W = 1080, 1920
frames = [np.random.randint(0, 255, (H, W, 3), dtype=np.uint8) for _ in range(64)]
dev = torch.empty((64, H, W, 3), dtype=torch.uint8, device="cuda")
for flags in (0, 2):   # default, cudaHostRegisterMapped(2)
    t0 = time.perf_counter()
    for f in frames:
        rc = rt.cudaHostRegister(f.ctypes.data, f.nbytes, flags)
        assert int(rc) == 0, rc
    t1 = time.perf_counter()
    for j, f in enumerate(frames):
        dev[j].copy_(torch.from_numpy(f), non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for f in frames:
        rt.cudaHostUnregister(f.ctypes.data)
    t3 = time.perf_counter()
    n = len(frames)
    print(f"flags={flags}: register {1e6 * (t1 - t0) / n:.0f} us/frame, H2D from registered {1e6 * (t2 - t1) / n:.0f} us/frame "
          f"({n * frames[0].nbytes / (t2 - t1) / 1e9:.1f} GB/s), unregister {1e6 * (t3 - t2) / n:.0f} us/frame")
# pageable H2D for comparison
t0 = time.perf_counter()
for j, f in enumerate(frames):
    dev[j].copy_(torch.from_numpy(f), non_blocking=True)
torch.cuda.synchronize()
print(f"pageable H2D {1e6 * (time.perf_counter() - t0) / len(frames):.0f} us/frame")
# single-thread and multi-thread memcpy into pinned staging
stage = torch.empty((64, H, W, 3), dtype=torch.uint8, pin_memory=True).numpy()
t0 = time.perf_counter()
for j, f in enumerate(frames):
    np.copyto(stage[j], f)
dt = time.perf_counter() - t0
print(f"1-thread copy into pinned: {1e6 * dt / 64:.0f} us/frame ({64 * frames[0].nbytes / dt / 1e9:.1f} GB/s)")
from concurrent.futures import ThreadPoolExecutor
for nt in (4, 8, 16, 32):
    with ThreadPoolExecutor(nt) as pool:
        t0 = time.perf_counter()
        for rep in range(4):
            list(pool.map(lambda j: np.copyto(stage[j], frames[j]), range(64)))
        dt = (time.perf_counter() - t0) / 4
    print(f"{nt}-thread copy into pinned: {1e6 * dt / 64:.0f} us/frame ({64 * frames[0].nbytes / dt / 1e9:.1f} GB/s)")
