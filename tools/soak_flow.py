"""Randomised parity soak of the sparse keypoint cadence: GPU path (CoordinateModel over PropagatedPath) against
the oracle frame loop run with the reference's own cv2 calls (oracle/pipeline.py, library_calls=True), on
rendered clips with random sizes, cadences, pan speeds, blanked / thinned head heatmaps, black frames and
calibration.  Run on the GPU box:  python tools/soak_flow.py [--clips 30] [--seed 0]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=30)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", type=int, default=-1, help="re-run a single clip of the sequence (same random draws, nothing else rendered)")
    args = ap.parse_args()
    import torch
    from eagle_b200 import synthetic
    from eagle_b200.coordinate_model import CoordinateModel
    from eagle_b200.engine import GeometryEngine
    from oracle import pipeline

    eng = GeometryEngine("cuda:0")
    rng = np.random.default_rng(args.seed)
    tot = {"clips": 0, "frames": 0, "dict_mismatch_clips": 0, "index_errors_agree": 0, "fallback_frames": 0, "repaired_chains": 0,
           "first_frame_rescues": 0, "pieces": 0}
    t0 = time.time()
    for ci in range(args.clips):
        W, H = [(640, 360), (854, 480), (1280, 720), (960, 540)][int(rng.integers(0, 4))]
        n = int(rng.integers(6, 29))
        fps, nh, nk = [(8, 1, 2), (24, 1, 3), (25, 1, 3), (6, 3, 1), (5, 1, 5), (12, 2, 4), (30, 1, 2)][int(rng.integers(0, 7))]
        cal = bool(rng.uniform() < 0.35)
        clip_seed, pan, hide = int(rng.integers(1 << 30)), float(rng.uniform(0.5, 7.0)), float(rng.uniform(0.05, 0.5))
        k = max(1, int(fps / max(1, nk)))
        if args.only >= 0 and ci != args.only:  # consume the same draws, skip the work
            for i in range(0, n, k):
                rng.uniform()
            for i in range(n):
                rng.uniform()
            rng.choice([2048, k, 2 * k, 3 * k])
            continue
        clip = synthetic.make_flow_clip(n, W, H, seed=clip_seed, pan_px=pan, hide_prob=hide)
        for i in range(0, n, k):
            u = rng.uniform()
            if u < 0.12:
                clip["heatmaps"][i] = 0.01
            elif u < 0.22:
                clip["heatmaps"][i, 3:] = 0.01
        for i in range(n):
            if rng.uniform() < 0.04:
                clip["frames"][i] = 0
        frames = list(clip["frames"])
        hm = torch.from_numpy(clip["heatmaps"]).cuda()
        x = eng.preprocess(torch.from_numpy(np.ascontiguousarray(clip["frames"])).cuda())
        sig = x[:, :, ::37, ::41].double().sum(dim=(1, 2, 3))
        black = [i for i in range(n) if not clip["frames"][i].any()]

        def net(t, sig=sig, hm=hm, black=black):
            out = []
            for v in t[:, :, ::37, ::41].double().sum(dim=(1, 2, 3)):
                m = torch.nonzero(sig == v).flatten().tolist()
                out.append(m[0])  # identical (black) frames share a signature; they also share blank-equivalent handling below
            return hm[out]

        # black frames are indistinguishable by content: give them all the same (blank) heatmap so the stand-in is well defined
        for i in black:
            clip["heatmaps"][i] = 0.01
        hm.copy_(torch.from_numpy(clip["heatmaps"]))
        objs = iter(clip["objects"])
        model = CoordinateModel(keypoint_model=net, detect_objects=lambda fr: next(objs), chunk=8)
        model.always_propagate = True
        model.piece_frames = int(rng.choice([2048, k, 2 * k, 3 * k]))
        err_g = err_o = None
        try:
            got = model.get_coordinates(frames, fps=fps, num_homography=nh, num_keypoint_detection=nk, verbose=False, calibration=cal)
        except IndexError as ex:
            err_g = str(ex)
        try:
            want = pipeline.get_coordinates_propagated(frames, clip["heatmaps"], clip["objects"], fps, nh, nk, calibration=cal, library_calls=True)
        except IndexError as ex:
            err_o = str(ex)
        tot["clips"] += 1; tot["frames"] += n
        if err_g or err_o:
            ok = bool(err_g) and bool(err_o)
            tot["index_errors_agree"] += ok
        else:
            ok = json.dumps(got, default=float) == json.dumps(want, default=float)
            st = model.last_stats
            tot["fallback_frames"] += st["fallback_frames"]; tot["repaired_chains"] += st["repaired_chains"]
            tot["first_frame_rescues"] += int(st["first_frame_rescue"]); tot["pieces"] += st["pieces"]
        if not ok:
            tot["dict_mismatch_clips"] += 1
            bad = None if (err_g or err_o) else next(i for i in want if json.dumps(got[i], default=float) != json.dumps(want[i], default=float))
            print(f"MISMATCH clip {ci}: {W}x{H} n={n} fps={fps} nh={nh} nk={nk} cal={cal} piece={model.piece_frames} first bad frame {bad} "
                  f"errors gpu={err_g} oracle={err_o}", flush=True)
            if args.only >= 0 and not (err_g or err_o):
                for i in want:
                    keys = [key for key in want[i] if json.dumps(got[i][key], default=float) != json.dumps(want[i][key], default=float)]
                    if keys:
                        print(f"  frame {i}: differing {keys}")
                        if "Keypoints" in keys:
                            print("    gpu   ", json.dumps(got[i]["Keypoints"], default=float))
                            print("    oracle", json.dumps(want[i]["Keypoints"], default=float))
                        if "Boundaries" in keys:
                            print("    gpu   ", got[i]["Boundaries"], "oracle", want[i]["Boundaries"])
    tot["seconds"] = round(time.time() - t0, 1)
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
