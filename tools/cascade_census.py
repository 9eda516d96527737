#!/usr/bin/env python
"""Census of the RANSAC -> RHO -> LMEDS cascade of coordinate_model.py:354-357 on live cv2 (CPU tool).

The CUDA path implements the RANSAC leg and reports EGL_FIT_NO_MODEL when it returns None.  That equals the
reference only if the later legs never return a homography for a set RANSAC gave up on.  This tool drives
`cv2.findHomography` with the reference's arguments over randomised *non-degenerate* low-inlier point sets drawn
from the pitch landmark table and counts, per family, how often RANSAC returns None and how often RHO or LMEDS
then rescue the frame.

    python tools/cascade_census.py [--sets 120000] [--workers 8] [--out profiles/r2_cascade_census.json]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FAMILIES = ("camera_outliers", "two_lines", "garbage", "near_collinear", "duplicates", "partial_flip", "one_line_plus_k",
            "tiny_spread")


def _world():
    from eagle_b200.pitch import WORLD_XYZ as LANDMARK_WORLD, OFF_PLANE
    on = [i for i in range(57) if i not in OFF_PLANE]
    return np.asarray(LANDMARK_WORLD, dtype=np.float64)[:, :2], on


def _camera(rng, W=1920, H=1080):
    """pitch -> image homography of a plausible broadcast camera (inverted later by the fit)."""
    f = rng.uniform(900, 4000)
    pan, tilt = rng.uniform(-0.7, 0.7), rng.uniform(0.15, 0.6)
    cx, cy, cz = rng.uniform(30, 75), rng.uniform(-60, -20), rng.uniform(8, 35)
    Rz = np.array([[np.cos(pan), -np.sin(pan), 0], [np.sin(pan), np.cos(pan), 0], [0, 0, 1]])
    t = np.pi / 2 + tilt
    Rx = np.array([[1, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]])
    R = Rx @ Rz.T
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    P = K @ np.c_[R, -R @ np.array([cx, cy, cz])]
    return P[:, [0, 1, 3]]


def make_set(rng, family, world, on, return_sel=False):
    """-> (img (n,2) float32, wor (n,2) float32[, landmark indices]); integer pixel positions like the reference's keypoints."""
    lines_x = {}
    lines_y = {}
    for i in on:
        lines_x.setdefault(round(world[i, 0], 3), []).append(i)
        lines_y.setdefault(round(world[i, 1], 3), []).append(i)
    lines = [v for v in list(lines_x.values()) + list(lines_y.values()) if len(v) >= 3]
    n = int(rng.integers(4, 13))
    if family == "two_lines":
        a, b = rng.choice(len(lines), 2, replace=False)
        pool = sorted(set(lines[a]) | set(lines[b]))
        n = min(n, len(pool))
        sel = np.sort(rng.choice(pool, n, replace=False))
    elif family == "one_line_plus_k":
        a = lines[int(rng.integers(len(lines)))]
        k = int(rng.integers(1, 3))
        rest = [i for i in on if i not in a]
        m = min(len(a), max(3, n - k))
        sel = np.sort(np.r_[rng.choice(a, m, replace=False), rng.choice(rest, k, replace=False)])
    else:
        sel = np.sort(rng.choice(on, n, replace=False))
    n = len(sel)
    wor = world[sel]
    Hc = _camera(rng)
    p = (Hc @ np.c_[wor, np.ones(n)].T).T
    img = p[:, :2] / p[:, 2:3]
    if family in ("camera_outliers", "two_lines", "one_line_plus_k"):
        img = img + rng.normal(0, rng.uniform(0, 3), img.shape)
        bad = rng.random(n) < rng.uniform(0.3, 0.8)
        img[bad] = rng.uniform(0, (1920, 1080), (int(bad.sum()), 2))
    elif family == "garbage":
        img = rng.uniform(0, (1920, 1080), (n, 2))
    elif family == "near_collinear":
        t = rng.uniform(0, 1, n)
        a, b = rng.uniform(0, (1920, 1080), 2), rng.uniform(0, (1920, 1080), 2)
        img = a + t[:, None] * (b - a) + rng.normal(0, rng.uniform(0.3, 3.0), (n, 2))
    elif family == "duplicates":
        img = img + rng.normal(0, 2, img.shape)
        k = int(rng.integers(1, max(2, n - 2)))
        img[rng.choice(n, k, replace=False)] = img[int(rng.integers(n))] + rng.integers(0, 2, (k, 2))
    elif family == "partial_flip":
        img = img + rng.normal(0, 1.5, img.shape)
        flip = rng.random(n) < 0.5
        img[flip, 0] = 1920 - img[flip, 0]
    elif family == "tiny_spread":
        img = rng.uniform(0, (1920, 1080), 2) + rng.integers(-3, 4, (n, 2))
    img = np.floor(np.clip(img, -4000, 6000)).astype(np.float32)
    return (img, wor.astype(np.float32), sel) if return_sel else (img, wor.astype(np.float32))


def worker(args):
    seed, count = args
    import cv2
    cv2.setNumThreads(1)
    world, on = _world()
    rng = np.random.default_rng(seed)
    out = {f: {"sets": 0, "ransac_none": 0, "rho_rescues": 0, "lmeds_rescues": 0, "ransac_ok_le5_inliers": 0} for f in FAMILIES}
    examples = []
    for t in range(count):
        fam = FAMILIES[t % len(FAMILIES)]
        img, wor = make_set(rng, fam, world, on)
        # non-degenerate by construction of the question: at least 4 distinct pixels, not all on one image line
        if len(np.unique(img, axis=0)) < 4:
            continue
        c = img - img.mean(0)
        if np.linalg.matrix_rank(c, tol=1e-3) < 2:
            continue
        st = out[fam]
        st["sets"] += 1
        H, mask = cv2.findHomography(img, wor, cv2.RANSAC, 5.0)
        if H is not None:
            st["ransac_ok_le5_inliers"] += int(mask.sum() <= 5)
            continue
        st["ransac_none"] += 1
        Hr, _ = cv2.findHomography(img, wor, cv2.RHO, None)
        if Hr is not None:
            st["rho_rescues"] += 1
            if len(examples) < 4:
                examples.append({"family": fam, "method": "RHO", "img": img.tolist(), "wor": wor.tolist()})
            continue
        Hl, _ = cv2.findHomography(img, wor, cv2.LMEDS, None)
        if Hl is not None:
            st["lmeds_rescues"] += 1
            if len(examples) < 4:
                examples.append({"family": fam, "method": "LMEDS", "img": img.tolist(), "wor": wor.tolist()})
    return out, examples


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=120000)
    ap.add_argument("--workers", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_cascade_census.json"))
    a = ap.parse_args()
    per = a.sets // a.workers
    with mp.get_context("fork").Pool(a.workers) as pool:
        parts = pool.map(worker, [(1000 + w, per) for w in range(a.workers)])
    import cv2
    total = {f: {k: sum(p[0][f][k] for p in parts) for k in parts[0][0][f]} for f in FAMILIES}
    examples = [e for p in parts for e in p[1]][:8]
    summary = {"cv2": cv2.__version__, "sets": sum(v["sets"] for v in total.values()),
               "ransac_none": sum(v["ransac_none"] for v in total.values()),
               "rho_rescues": sum(v["rho_rescues"] for v in total.values()),
               "lmeds_rescues": sum(v["lmeds_rescues"] for v in total.values()), "families": total, "examples": examples}
    json.dump(summary, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "examples"}, indent=1))


if __name__ == "__main__":
    main()
