"""Randomised component soaks on the GPU box: K1 at odd sizes, K2 at odd map shapes, K4 on wild homographies,
fixed-K fit vs the C mirror at varied N / outlier rates.  Prints one JSON summary."""
import json
import os
import sys
import warnings

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from eagle_b200 import _native as N  # noqa: E402
from eagle_b200 import synthetic  # noqa: E402
from eagle_b200.engine import GeometryEngine, KeypointSet  # noqa: E402
from eagle_b200.pitch import WORLD_XY_F32  # noqa: E402
from oracle import preprocess, project, ransac_f32  # noqa: E402

eng = GeometryEngine("cuda:0")
rng = np.random.default_rng(2024)
out = {}

# ---- K1: arbitrary sizes (unaligned rows -> non-bulk path, upsampling, strided views) -----------------
bad = 0; n = 0
sizes = [(480, 854), (481, 855), (1079, 1919), (1081, 1921), (2161, 3841), (100, 100), (2, 2), (17, 4001), (540, 960), (541, 961), (3240, 5760), (720, 1281)]
for (h, w) in sizes:
    fr = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    got = eng.preprocess(torch.from_numpy(fr).cuda()).cpu().numpy()
    for i in range(2):
        want = preprocess.preprocess_reference_calls(fr[i])
        n += 1
        if not np.max(np.abs(got[i] - want)) <= 1e-6:
            bad += 1; print("K1 mismatch", h, w, float(np.max(np.abs(got[i] - want))))
# padded rows / frames (row_stride > 3W): a view into a larger buffer
big = torch.from_numpy(rng.integers(0, 256, (3, 730, 1300, 3), dtype=np.uint8)).cuda()
view = big[:, 5:725, 10:1290, :]
got = eng.preprocess(view).cpu().numpy()
for i in range(3):
    want = preprocess.preprocess_reference_calls(np.ascontiguousarray(view[i].cpu().numpy()))
    n += 1
    if not np.max(np.abs(got[i] - want)) <= 1e-6:
        bad += 1; print("K1 strided mismatch", i)
out["K1_cases"] = n; out["K1_mismatches"] = bad

# ---- K2: odd map shapes and frame counts ----------------------------------------------------------------
bad = 0; n = 0
for (F, h, w) in [(1, 135, 240), (3, 68, 120), (2, 270, 480), (5, 1, 4), (7, 33, 44), (4, 540, 960), (300, 135, 240), (2, 135, 244)]:
    hm = torch.rand((F, 57, h, w), device="cuda")
    hm[0, 0] = 0.5                                            # constant map -> index 0
    kp = eng.decode(hm, 1920, 1080)
    ok = bool((kp.flat.long() == hm.view(F, 57, -1).argmax(2)).all()) and bool((kp.score == hm.view(F, 57, -1).amax(2)).all())
    n += 1; bad += not ok
    if not ok: print("K2 mismatch", F, h, w)
out["K2_cases"] = n; out["K2_mismatches"] = bad

# ---- K4: wild homographies (near-singular denominators, huge values, NaN) -------------------------------
bad = 0; n = 0
F, P = 400, 23
Hs = rng.normal(0, 1, (F, 9)) * np.array([1e-1, 1e-1, 30, 1e-2, 2e-1, 30, 1e-5, 1e-3, 1])
Hs[:, 8] = 1.0
Hs[5] = [1, 0, 0, 0, 1, 0, 0, 0, 1]            # identity: corners share y -> ZeroDivisionError -> None
Hs[6] = [0, 0, 0, 0, 0, 0, 0, 0, 0]            # w == 0 everywhere -> (0, 0)
Hs[7, 6:8] = [-1 / 960.0, 0]                   # denominator crosses zero inside the frame
Hs[8] = np.nan
pts = rng.uniform(0, 1920, (F, P, 2)).astype(np.float32)
cnt = rng.integers(0, P + 1, F).astype(np.int32)
pr = eng.project(torch.from_numpy(Hs).cuda(), torch.from_numpy(pts).cuda(), torch.from_numpy(cnt).cuda(), 1920, 1080)
cf = pr.coords.cpu().numpy(); ci = pr.coords_i.cpu().numpy(); ib = pr.in_bounds.cpu().numpy(); bd = pr.bounds.cpu().numpy()
for f in range(F):
    H = Hs[f].reshape(3, 3)
    k = int(cnt[f])
    if k:
        want = cv2.perspectiveTransform(pts[f:f + 1, :k], H)[0]
        wi = want.astype(int)
        n += 1
        same_f = np.array_equal(cf[f, :k], want) or np.array_equal(np.nan_to_num(cf[f, :k], nan=-7.0), np.nan_to_num(want, nan=-7.0))
        if not (same_f and np.array_equal(ci[f, :k], wi)):
            bad += 1; print("K4 point mismatch", f)
        wib = ~((wi[:, 0] < 0) | (wi[:, 0] > 105) | (wi[:, 1] < 0) | (wi[:, 1] > 68))
        if not np.array_equal(ib[f, :k].astype(bool), wib):
            bad += 1; print("K4 bounds-flag mismatch", f)
    wb = project.boundaries(1920, 1080, H)
    n += 1
    if wb[0] is None:
        ok = bool(np.isnan(bd[f]).all())
    else:
        ok = all((np.isnan(a) and np.isnan(b[0])) or a == b[0] for a, b in zip(bd[f], wb))
    if not ok:
        bad += 1; print("K4 boundary mismatch", f, bd[f], wb)
out["K4_cases"] = n; out["K4_mismatches"] = bad

# ---- fixed-K fit vs the independent C mirror: varied N and outlier rates --------------------------------
bad = 0; n = 0
on = [i for i in range(57) if i not in (0, 1, 24, 25)]
F, K = 96, 256
xy = np.zeros((F, 57, 2), np.int32); order = np.full((F, 64), 255, np.uint8); count = np.zeros((F, 2), np.int32)
sels = []
for f in range(F):
    w, h = [(1280, 720), (1920, 1080), (3840, 2160)][f % 3]
    cam = synthetic.sample_cameras(1, w, h, rng)[0]
    nn = int(rng.integers(4, 54)); sel = np.sort(rng.choice(on, nn, replace=False))
    px = synthetic.project_points(cam, WORLD_XY_F32[sel].astype(np.float64)) + rng.normal(0, 0.5, (nn, 2))
    no = int(nn * rng.uniform(0, 0.5)); oi = rng.choice(nn, no, replace=False); px[oi] = rng.uniform([0, 0], [w, h], (no, 2))
    xy[f, sel] = np.rint(px).astype(np.int32); order[f, :nn] = sel; count[f] = nn; sels.append(sel)
hyp = np.stack([np.stack([rng.choice(max(4, int(count[f, 0])), 4, replace=False) for _ in range(K)]) for f in range(F)]).astype(np.uint8)
kp = KeypointSet(torch.zeros((F, 57), dtype=torch.int32).cuda(), torch.zeros((F, 57)).cuda(), torch.from_numpy(xy).cuda(),
                 torch.from_numpy(order).cuda(), torch.from_numpy(count).cuda())
fit = eng.fit(kp, mode=N.FIT_FIXED_K, K=K, hyp=torch.from_numpy(hyp).cuda())
st = fit.status.cpu().numpy(); info = fit.info.cpu().numpy(); inl = fit.inlier_mask.cpu().numpy(); Hg = fit.H.cpu().numpy().reshape(-1, 3, 3)
worst = 0.0
for f in range(F):
    sel = sels[f]; img = xy[f, sel].astype(np.float32); wor = WORLD_XY_F32[sel]
    Ho, mo, r = ransac_f32.fit_fixedk(img, wor, hyp[f])
    n += 1
    if Ho is None:
        if st[f] == 0: bad += 1; print("fixed-K: oracle none, gpu ok", f, info[f])
        continue
    if st[f] != 0 or info[f, 2] != r["best_index"] or int(inl[f]) != sum(1 << int(c) for c, b in zip(sel, mo.ravel()) if b):
        bad += 1; print("fixed-K mismatch", f, len(sel), st[f], info[f].tolist(), r["best_index"])
    elif int(mo.sum()) >= 8:
        worst = max(worst, float(np.max(np.abs(Hg[f] - Ho) / np.abs(Ho))))
out["fixedK_cases"] = n; out["fixedK_mismatches"] = bad; out["fixedK_worst_rel_H_(>=8 inliers)"] = worst
print(json.dumps(out))
