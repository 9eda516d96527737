#!/usr/bin/env python
"""CoordinateModel.get_coordinates on 2250 pageable 1080p frames (stand-in network and detector, as bench.py's api_e2e)
for several upload thread counts and chunk sizes, in one process on one box: frames/s and time inside the upload calls."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eagle_b200 import synthetic
from eagle_b200.coordinate_model import CoordinateModel
F, H, W = 2250, 1080, 1920
dev = "cuda:0"
clip = synthetic.make_clip(75, W, H, seed=5, ghost_prob=0.05)
hm = torch.from_numpy(clip["heatmaps"]).to(dev)
hm = hm.repeat((F + 74) // 75, 1, 1, 1)[:F].contiguous()
objs_pool = clip["objects"]
host_pool = np.random.default_rng(0).integers(0, 256, (256, H, W, 3), dtype=np.uint8)
host_frames = [host_pool[i % 256] for i in range(F)]
state = {"i": 0, "h": 0}
def detector(_f):
    o = objs_pool[state["i"] % len(objs_pool)]; state["i"] += 1; return o
def network(x):
    n = x.shape[0]; s = state["h"] % F
    if s + n > F: s = 0
    state["h"] = s + n
    return hm[s:s + n]
out = []
for chunk in (75, 150):
    model = CoordinateModel(keypoint_model=network, detect_objects=detector, device=dev, chunk=chunk)
    model.network_batch = chunk
    for nt in (4, 6, 8, 10, 12, 16):
        model.copy_threads = nt
        if model._stream is not None:
            model._stream.copy_threads = nt
        def once():
            state["i"] = 0; state["h"] = 0
            return model.get_coordinates(host_frames, fps=25, num_homography=25, num_keypoint_detection=25, verbose=False)
        once()
        t0 = time.perf_counter(); reps = 3
        for _ in range(reps): res = once()
        dt = (time.perf_counter() - t0) / reps
        st = model.last_stats
        row = {"chunk": chunk, "threads": nt, "frames_per_s": F / dt, "upload_us_per_frame": st["upload_s"] / F * 1e6}
        print(json.dumps(row), flush=True); out.append(row)
