"""The kernels' scalar core (eagle_b200/csrc/geometry_core.cuh), compiled for the host, against the
oracle and the live cv2.  This is the same source the GPU executes; the -m gpu tests then check the
device build and the kernel orchestration around it."""
import json
import os
import warnings

import cv2
import numpy as np
import pytest

import hostcore
from eagle_b200 import synthetic
from eagle_b200.pitch import WORLD_XY_F32
from oracle import decode, homography, landmarks, synthesis

ON = [i for i in range(57) if i not in (0, 1, 24, 25)]


def random_case(rng, t):
    W, H = [(1280, 720), (1920, 1080), (3840, 2160)][t % 3]
    cam = synthetic.sample_cameras(1, W, H, rng)[0]
    n = int(rng.integers(5, 54)); sel = np.sort(rng.choice(ON, n, replace=False))
    px = synthetic.project_points(cam, WORLD_XY_F32[sel].astype(np.float64)) + rng.normal(0, rng.choice([0, .5, 2.]), (n, 2))
    no = int(n * rng.uniform(0, .5)); oi = rng.choice(n, no, replace=False)
    px[oi] = rng.uniform([0, 0], [W, H], (no, 2))
    return np.rint(px).astype(np.float32), WORLD_XY_F32[sel]


def test_cv2_compatible_fit_matches_live_cv2():
    rng = np.random.default_rng(5)
    worst = 0.0
    for t in range(150):
        img, wor = random_case(rng, t)
        Hc, mc = cv2.findHomography(img, wor, cv2.RANSAC, 5.0)
        st, Hh, mh, info = hostcore.fit_cv2(img, wor)
        if Hc is None:
            assert st != 0
            continue
        assert st == 0
        assert np.array_equal(mc.ravel(), mh), t
        if mc.sum() >= 6:
            worst = max(worst, float(np.max(np.abs(Hc - Hh) / np.abs(Hc))))
    assert worst < 1e-5, worst  # north-star tolerance is 1e-4


def test_minimal_solvers_agree_with_opencv_run_kernel():
    rng = np.random.default_rng(2)
    for t in range(100):
        img, wor = random_case(rng, t)
        idx = rng.choice(len(img), 4, replace=False)
        if not homography.check_subset(img[idx], wor[idx]):
            assert not hostcore.check_subset(img[idx], wor[idx])
            continue
        assert hostcore.check_subset(img[idx], wor[idx])
        Hk = homography.run_kernel(img[idx], wor[idx])
        ok64, H64 = hostcore.dlt4_f64(img[idx], wor[idx])
        assert ok64 and np.max(np.abs(H64 - Hk) / (np.abs(Hk) + 1e-12)) < 1e-7


def test_postprocess_matches_reference_decoded_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_small.npz"))
    hm = g["heatmaps"]; N, C, h, w = hm.shape
    post = json.loads(str(g["postprocessed_json"]))
    for key, want in post.items():
        wh, n = key.split(":"); W, H = (int(v) for v in wh.split("x")); n = int(n)
        flat = hm[n].reshape(C, -1).argmax(1); score = hm[n].reshape(C, -1).max(1)
        xy, order = hostcore.postprocess(flat, score, h, w, W, H)
        got = {landmarks.INDEX_TO_NAME[int(c)]: [int(xy[c, 0]), int(xy[c, 1])] for c in order}
        assert got == want and list(got) == list(want), key


def test_synthesis_matches_oracle():
    warnings.simplefilter("ignore")
    clip = synthetic.make_clip(24, 1920, 1080, seed=3, ghost_prob=0.05)
    added = 0
    for i in range(24):
        kp = decode.decode_frame(clip["heatmaps"][i], 1920, 1080)
        want = synthesis.synthesize(kp)
        xy = np.zeros((57, 2), np.int32); order = []
        for name, (x, y) in kp.items():
            c = landmarks.NAME_TO_INDEX[name]; xy[c] = (x, y); order.append(c)
        xy2, order2 = hostcore.synthesize(xy, np.array(order, np.uint8))
        got = {landmarks.INDEX_TO_NAME[int(c)]: (int(xy2[c, 0]), int(xy2[c, 1])) for c in order2}
        assert list(got) == list(want)
        assert all(tuple(want[k]) == got[k] for k in want)
        added += len(want) - len(kp)
    assert added > 20  # the fixture does exercise the synthesis


def test_projection_matches_cv2():
    rng = np.random.default_rng(0)
    H = np.linalg.inv(synthetic.sample_cameras(1, 1920, 1080, rng)[0]); H /= H[2, 2]
    pts = rng.uniform(0, 1920, (2000, 2)).astype(np.float32)
    of, oi = hostcore.project(H, pts)
    ref = cv2.perspectiveTransform(pts[None], H)[0]
    assert np.array_equal(of, ref) and np.array_equal(oi, ref.astype(int))


def test_seeded_generator_is_distinct_and_in_range():
    for N in (4, 5, 17, 53):
        for h in range(200):
            idx = hostcore.seeded_subset(123, 7, 4096, h, N)
            assert len(set(idx.tolist())) == 4 and idx.min() >= 0 and idx.max() < N


def test_fixed_k_core_equals_independent_c_mirror():
    """The kernel's FP32 hypothesis arithmetic (host build) vs oracle/ransac_f32.c, written separately
    from the specification: same winner, same normalised-test count, same H to rounding."""
    from oracle import ransac_f32
    xy, valid, flags, cams = synthetic.stress_point_sets(10, 1920, 1080, seed=5)
    K = 768
    for f in range(10):
        img = xy[f, ON].astype(np.float32); wor = WORLD_XY_F32[ON]
        table = ransac_f32.seeded_table(77, f, K, 53)
        assert all(np.array_equal(table[h], hostcore.seeded_subset(77, f, K, h, 53)) for h in range(0, K, 29))
        st, Hb, m, info = hostcore.fixedk_stage(img, wor, K, None, seed=77, frame=f)
        r = ransac_f32.fixedk_frame(img, wor, table)
        assert st == 0 and info[2] == r["best_index"]
        assert np.max(np.abs(ransac_f32.denormalise(r["h"], r["norm"], 5.0) - Hb) / np.abs(Hb)) < 1e-12
        H, mask, _ = ransac_f32.fit_fixedk(img, wor, table)
        Hr, fm, n = hostcore.refit(Hb, img, wor, m)
        assert np.array_equal(mask.ravel(), fm) and np.max(np.abs(H - Hr) / np.abs(Hr)) < 1e-6
        assert not any(fm[k] for k, c in enumerate(ON) if flags[f, c])


def test_postprocess_randomised_collisions_and_ties():
    """Many channels landing on few pixels with scores from a small discrete set: exercises the duplicate
    arbitration (best score wins; exact ties -> later label, earlier dict slot) against the oracle."""
    rng = np.random.default_rng(4)
    h, w = 12, 20
    vals = np.array([0.009, 0.01, 0.0100001, 0.2, 0.29999998, 0.3, 0.30000001, 0.5, 0.5, 0.9, 1.0], np.float32)
    for t in range(300):
        pix = rng.choice(h * w, size=rng.integers(1, 6), replace=False)
        flat = rng.choice(pix, size=57)
        score = rng.choice(vals, size=57)
        W, H = [(1280, 720), (1920, 1080), (3840, 2160), (854, 480)][t % 4]
        kps = [(i, (flat[i] % w) / (w - 1), (flat[i] // w) / (h - 1), float(score[i])) for i in range(57) if float(score[i]) > 0.01]
        want = decode.postprocess(kps, W, H, 0.3)
        xy, order = hostcore.postprocess(flat, score, h, w, W, H)
        got = {landmarks.INDEX_TO_NAME[int(c)]: (int(xy[c, 0]), int(xy[c, 1])) for c in order}
        assert list(got) == list(want) and got == {k: tuple(v) for k, v in want.items()}, t


# ---- keypoint propagation: the kernels' scalar flow code (csrc/flow_core.cuh) on the host ------------------
def test_flow_core_colour_and_pyramid_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    v = np.unique(np.r_[np.arange(0, 256, 5), 254, 255]).astype(np.uint8)
    c = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(len(v), -1, 3)
    g, h = hostcore.gray_hue(c)
    assert np.array_equal(g, cv2.cvtColor(c, cv2.COLOR_BGR2GRAY)) and np.array_equal(h, cv2.cvtColor(c, cv2.COLOR_BGR2HSV)[..., 0])
    for shape in [(37, 53), (135, 240), (64, 65), (33, 18)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        lv = hostcore.build_pyramid(img, 2)
        want = [img]
        for _ in range(2):
            nxt = cv2.pyrDown(want[-1])
            if nxt.shape[0] <= 15 or nxt.shape[1] <= 15:
                break
            want.append(nxt)
        assert len(lv) == len(want) and all(np.array_equal(a, b) for a, b in zip(lv, want))
    for n in list(range(0, 20)) + [31, 57, 64, 100]:
        a = (rng.uniform(0, 30, n) ** 2).astype(np.float32)
        assert hostcore.pairwise_sum(a) == (a.sum() if n else np.float32(0))


@pytest.mark.parametrize("seed", [2, 3, 4])
def test_flow_core_tracker_bit_exact_vs_cv2(seed):
    """The CUDA tracker's scalar statement (lk_track_point, also the EGL_TRACK_VARIANT=1 kernel), compiled for
    the host, against live cv2.calcOpticalFlowPyrLK: status and float32 positions bit for bit."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    tracked = 0
    for trial in range(10):
        h, w = [(270, 480), (135, 240), (100, 70), (360, 640), (40, 50)][trial % 5]
        cell = [2, 4, 8, 16][trial % 4]
        base = rng.integers(0, 256, (h // cell + 2, w // cell + 2), dtype=np.uint8)
        g1 = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
        if trial % 3 == 0:
            g1 = (g1.astype(np.int32) + rng.integers(-40, 40, g1.shape)).clip(0, 255).astype(np.uint8)
        if trial == 7:
            g1 = rng.integers(0, 256, (h, w), dtype=np.uint8)   # white noise: every float sum rounds
        ang = rng.uniform(-0.03, 0.03); tx, ty = rng.uniform(-6, 6, 2)
        M = np.float32([[np.cos(ang), np.sin(ang), tx], [-np.sin(ang), np.cos(ang), ty]])
        g2 = cv2.warpAffine(g1, M, (w, h), borderMode=cv2.BORDER_REFLECT)
        pts = rng.uniform([-3, -3], [w + 2, h + 2], (40, 2)).astype(np.float32)
        if trial % 2 == 0:
            pts = np.trunc(pts)
        ref, st, _ = cv2.calcOpticalFlowPyrLK(g1, g2, pts, None, **lk)
        out, s = hostcore.track(g1, g2, pts)
        assert np.array_equal(s, st[:, 0]), trial
        ok = st[:, 0] == 1
        assert np.array_equal(out[ok].view(np.int32), ref[ok].view(np.int32)), trial
        tracked += int(ok.sum())
    assert tracked > 250


def test_few_inlier_fits_follow_cv2_eigenvalue_threshold():
    """With <= 9 inliers J^T J has eigenvalues below the threshold under which cv::solve(DECOMP_EIG) drops them;
    the scalar refit follows it there (eig_threshold_solve9).  Regression case: frame 8 of a soak clip, where a
    black network frame made the reference fit a homography to meaningless flowed positions -- cv2 keeps 6 inliers
    and so must the kernels' scalar code; then a sweep over hard synthetic point sets."""
    cv2 = pytest.importorskip("cv2")
    ip = np.array([[57, 332], [98, 393], [461, 89], [854, 258], [895, 40], [644, 328], [333, 147], [698, 381], [450, 163], [566, 357],
                   [609, 345], [527, 199], [199, 393], [156, 356], [151, 340], [182, 411], [163, 306], [863, 127], [898, 379], [846, 76],
                   [646, 397], [136, 377]], np.float32)
    wp = np.array([[5.5, 24.84], [5.5, 43.16], [16.5, 13.84], [16.5, 54.16], [0.0, 54.16], [52.5, 0.0], [88.5, 13.84], [105.0, 0.0],
                   [61.31, 36.46], [43.69, 36.46], [61.31, 31.54], [43.69, 31.54], [58.97, 40.47], [58.97, 27.53], [52.5, 43.15],
                   [52.5, 24.85], [20.15, 34.0], [19.99, 35.7], [19.99, 32.3], [11.0, 34.0], [16.5, 34.0], [52.5, 34.0]], np.float32)
    H, m = cv2.findHomography(ip, wp, cv2.RANSAC, 5.0)
    st, Hh, mh, info = hostcore.fit_cv2(ip, wp)
    assert st == 0 and int(m.sum()) == 6 and np.array_equal(m.ravel(), mh)
    assert np.max(np.abs(Hh - H) / np.abs(H)) < 1e-6

    from eagle_b200 import synthetic
    from eagle_b200.pitch import OFF_PLANE, WORLD_XYZ
    on = np.array([i for i in range(57) if i not in OFF_PLANE])
    rng = np.random.default_rng(0)
    fits = mask_diff = h_bad = 0
    for t in range(400):
        W, Himg = [(1280, 720), (1920, 1080), (960, 540)][t % 3]
        cam = synthetic.sample_cameras(1, W, Himg, rng)[0]
        px, vis = synthetic.landmark_pixels(cam, W, Himg)
        sel = on[vis[on]]
        if len(sel) < 6:
            continue
        if t % 2 == 0:   # a handful of true inliers among gross outliers
            good = rng.choice(sel, min(int(rng.integers(6, 9)), len(sel)), replace=False)
            bad = rng.choice(np.setdiff1d(on, good), int(rng.integers(5, 20)), replace=False)
            pts = {int(c): px[c] + rng.normal(0, 0.7, 2) for c in good}
            pts.update({int(c): rng.uniform([0, 0], [W, Himg]) for c in bad})
        else:            # nothing but scattered positions
            chs = rng.choice(on, int(rng.integers(8, 30)), replace=False)
            base = rng.uniform([0, 0], [W, Himg])
            pts = {int(c): base + rng.normal(0, rng.uniform(20, 300), 2) for c in chs}
        chs = sorted(pts)
        a = np.rint(np.array([pts[c] for c in chs])).astype(np.float32); b = WORLD_XYZ[chs, :2].astype(np.float32)
        H, m = cv2.findHomography(a, b, cv2.RANSAC, 5.0)
        st, Hh, mh, info = hostcore.fit_cv2(a, b)
        if H is None or int(m.sum()) < 6:
            continue
        fits += 1
        if st != 0 or not np.array_equal(m.ravel(), mh):
            mask_diff += 1
        elif np.max(np.abs(Hh - H) / np.abs(H)) > 1e-4:
            h_bad += 1
    assert fits > 100 and mask_diff == 0 and h_bad == 0, (fits, mask_diff, h_bad)
