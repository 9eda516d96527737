#!/usr/bin/env python
"""CoordinateModel.get_coordinates on 2250 pageable 1080p frames (stand-in network and detector, as bench.py's api_e2e):
one upload call at a time against two in flight (DenseStream.uploads_in_flight), for several thread counts per call.
One JSON line per configuration; the dict of every configuration is compared with the first one's."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.coordinate_model import CoordinateModel  # noqa: E402

F, H, W = 2250, 1080, 1920
dev = "cuda:0"
clip = synthetic.make_clip(75, W, H, seed=5, ghost_prob=0.05)
hm = torch.from_numpy(clip["heatmaps"]).to(dev)
hm = hm.repeat((F + 74) // 75, 1, 1, 1)[:F].contiguous()
objs_pool = clip["objects"]
base = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
host_pool = np.stack([np.bitwise_xor(base, np.uint8(i)) for i in range(256)])
host_frames = [host_pool[i % 256] for i in range(F)]
state = {"i": 0, "h": 0}


def detector(_f):
    o = objs_pool[state["i"] % len(objs_pool)]
    state["i"] += 1
    return o


def network(x):
    n = x.shape[0]
    s = state["h"] % F
    if s + n > F:
        s = 0
    state["h"] = s + n
    return hm[s:s + n]


want = None
for chunk, inflight, nt in ((150, 1, 8), (150, 2, 4), (150, 2, 5), (150, 2, 6), (150, 2, 8), (75, 2, 5), (150, 1, 8)):
    model = CoordinateModel(keypoint_model=network, detect_objects=detector, device=dev, chunk=chunk)
    model.network_batch = chunk
    model.copy_threads = nt
    model.uploads_in_flight = inflight

    def once():
        state["i"] = 0
        state["h"] = 0
        return model.get_coordinates(host_frames, fps=25, num_homography=25, num_keypoint_detection=25, verbose=False)

    res = once()
    got = json.dumps(res, default=float)
    if want is None:
        want = got
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        once()
    dt = (time.perf_counter() - t0) / reps
    st = model.last_stats
    print(json.dumps({"chunk": chunk, "uploads_in_flight": inflight, "threads_per_call": nt, "frames_per_s": round(F / dt, 1),
                      "h2d_GBps": round(F * H * W * 3 / dt / 1e9, 2), "upload_us_per_frame_summed_over_calls": round(st["upload_s"] / F * 1e6, 1),
                      "dict_identical_to_first": got == want}), flush=True)
    model._stream.close()
    del model
