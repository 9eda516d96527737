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

# pageable H2D issued from several threads, one stream each (the driver stages through its own bounce buffers)
import threading
frames2 = [np.random.randint(0, 255, (H, W, 3), dtype=np.uint8) for _ in range(128)]
dev2 = torch.empty((128, H, W, 3), dtype=torch.uint8, device="cuda")
for nt in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream() for _ in range(nt)]

    def work(k):
        with torch.cuda.stream(streams[k]):
            for j in range(k, 128, nt):
                dev2[j].copy_(torch.from_numpy(frames2[j]), non_blocking=True)
            streams[k].synchronize()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(nt)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"pageable H2D from {nt} thread(s)/stream(s): {1e6 * dt / 128:.0f} us/frame ({128 * frames2[0].nbytes / dt / 1e9:.1f} GB/s)")

# staging through a SMALL page-locked ring (cache-resident bounce buffers): each thread copies one 1 MB slice at a time
# into its own pinned slot and immediately issues the H2D of that slice
SL = 1 << 20
for nt in (4, 8, 16):
    slots = [torch.empty((4, SL), dtype=torch.uint8, pin_memory=True) for _ in range(nt)]
    streams = [torch.cuda.Stream() for _ in range(nt)]
    flat_dev = dev2.view(128, -1)
    nb = frames2[0].nbytes

    def work2(k):
        evs = [None] * 4
        q = 0
        with torch.cuda.stream(streams[k]):
            for j in range(k, 128, nt):
                src = frames2[j].reshape(-1)
                for o in range(0, nb, SL):
                    n = min(SL, nb - o)
                    if evs[q] is not None:
                        evs[q].synchronize()
                    np.copyto(slots[k][q].numpy()[:n], src[o:o + n])
                    flat_dev[j, o:o + n].copy_(slots[k][q][:n], non_blocking=True)
                    evs[q] = streams[k].record_event()
                    q = (q + 1) & 3
            streams[k].synchronize()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work2, args=(k,)) for k in range(nt)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"1 MB pinned ring, {nt} threads: {1e6 * dt / 128:.0f} us/frame ({128 * nb / dt / 1e9:.1f} GB/s)")
