"""Randomised parity soak (run on the GPU box): many synthetic frames at mixed resolutions and outlier
rates through the CUDA path and through the oracle (live cv2); reports every disagreement."""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.coordinate_model import GeometryPath  # noqa: E402
from oracle import pipeline  # noqa: E402

n_clips = int(os.environ.get("CLIPS", 24)); per = int(os.environ.get("FRAMES", 64))
path = GeometryPath("cuda:0")
bad_dict = bad_mask = n_fit = 0
worst_h = 0.0
worst_by = {}   # inlier-count bucket -> (frames, frames with rel H error > 1e-4, worst)
bad_dict_inl = []
t0 = time.time()
for c in range(n_clips):
    w, h = [(1280, 720), (1920, 1080), (3840, 2160), (854, 480)][c % 4]
    clip = synthetic.make_clip(per, w, h, seed=1000 + c, ghost_prob=[0.0, 0.05, 0.15, 0.3][(c // 4) % 4], hide_prob=[0.1, 0.3, 0.5][c % 3])
    trace = []
    want = pipeline.get_coordinates(clip["heatmaps"], clip["objects"], w, h, trace=trace)
    hm = torch.from_numpy(clip["heatmaps"]).cuda()
    got = path.run(hm, clip["objects"], w, h, fps=1)
    for i in range(per):
        if json.dumps(got[i], default=float, sort_keys=True) != json.dumps(want[i], default=float, sort_keys=True):
            bad_dict += 1
            bad_dict_inl.append(int(trace[i]["mask"].sum()) if trace[i]["mask"] is not None else -1)
            print(f"clip {c} frame {i}: dict differs ({w}x{h}), cv2 inliers {bad_dict_inl[-1]}")
    foot, count = synthetic.objects_to_arrays(clip["objects"], 23)
    kp, fit, hi, at, pr = path.run_device(hm, torch.from_numpy(foot).cuda(), torch.from_numpy(count).cuda(), w, h)
    Hs = fit.H.cpu().numpy().reshape(-1, 3, 3); inl = fit.inlier_mask.cpu().numpy(); st = fit.status.cpu().numpy()
    from eagle_b200.pitch import LANDMARK_INDEX
    for i, t in enumerate(trace):
        if t["H"] is None:
            bad_mask += st[i] == 0
            continue
        n_fit += 1
        chans = [LANDMARK_INDEX[n] for n in t["used_labels"]]
        wm = sum(1 << ch for ch, m in zip(chans, t["mask"].ravel()) if m)
        if st[i] != 0 or int(inl[i]) != wm:
            bad_mask += 1
            print(f"clip {c} frame {i}: mask/status differs")
        else:
            ninl = int(t["mask"].sum())
            rel = float(np.max(np.abs(Hs[i] - t["H"]) / np.abs(t["H"])))
            b = "<=5" if ninl <= 5 else ("6-7" if ninl <= 7 else ("8-11" if ninl <= 11 else ">=12"))
            fr, over, wst = worst_by.get(b, (0, 0, 0.0))
            worst_by[b] = (fr + 1, over + (rel > 1e-4), max(wst, rel))
            if ninl >= 6:
                worst_h = max(worst_h, rel)
print(json.dumps({"frames": n_clips * per, "fits": n_fit, "dict_mismatches": bad_dict, "mask_or_status_mismatches": int(bad_mask),
                  "worst_rel_H_error_(>=6 inliers)": worst_h,
                  "by_cv2_inlier_count_(frames, frames_over_1e-4, worst_rel)": worst_by, "dict_mismatch_inlier_counts": bad_dict_inl, "seconds": round(time.time() - t0, 1)}))
