"""CPU tests: the keypoint-propagation oracle (oracle/optflow.py) against the live libraries it restates
(cv2 4.13 / numpy 2.3 in this image) and, where /root/reference is mounted, against the reference's own
``calculate_optical_flow`` / ``calibrate_keypoints`` / ``get_coordinates``."""
import json

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from eagle_b200 import synthetic as S
from oracle import decode as D, optflow as O, pipeline as P, ref_harness as RH, synthesis as SY

LK = dict(winSize=(15, 15), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
needs_reference = pytest.mark.skipif(not RH.reference_available(), reason="/root/reference not mounted")


def _textured(h, w, rng, cell):
    base = rng.integers(0, 256, (h // cell + 2, w // cell + 2, 3), dtype=np.uint8)
    return cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)


def test_colour_pyramid_and_derivatives_match_cv2():
    rng = np.random.default_rng(0)
    for shape in [(37, 53), (270, 480), (64, 65), (15, 16)]:
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
        assert np.array_equal(g, O.gray_restated(img))
        assert np.array_equal(hsv[..., 0], O.hue_restated(img))
        assert np.array_equal(hsv[..., 2], O.value_restated(img))
        assert np.array_equal(cv2.pyrDown(g), O.pyr_down_restated(g))
        ix, iy = O.scharr_restated(g)
        assert np.array_equal(cv2.Scharr(g, cv2.CV_16S, 1, 0), ix) and np.array_equal(cv2.Scharr(g, cv2.CV_16S, 0, 1), iy)
    # every colour whose components are multiples of 3, plus the 0/255 extremes
    v = np.unique(np.r_[np.arange(0, 256, 3), 254, 255]).astype(np.uint8)
    c = np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(len(v), -1, 3)
    assert np.array_equal(cv2.cvtColor(c, cv2.COLOR_BGR2HSV)[..., 0], O.hue_restated(c))
    assert np.array_equal(cv2.cvtColor(c, cv2.COLOR_BGR2GRAY), O.gray_restated(c))


def test_numpy_float32_reductions_restated():
    rng = np.random.default_rng(1)
    for n in list(range(1, 24)) + [31, 32, 33, 57, 64]:
        for _ in range(5):
            new = rng.uniform(0, 900, (n, 2)).astype(np.float32)
            prev = (new + rng.normal(0, 3, (n, 2))).astype(np.float32)
            mv = np.linalg.norm(new - prev, axis=1)
            move, mean, std = O.move_stats_restated(new, prev)
            ref_std = np.std(mv) + 1e-6
            assert np.array_equal(mv, move) and mean == np.mean(mv) and std == ref_std and type(ref_std) is np.float32


@pytest.mark.parametrize("seed", [2, 3])
def test_lk_tracker_bit_exact_vs_cv2(seed):
    rng = np.random.default_rng(seed)
    total = 0
    for trial in range(6):
        h, w = [(270, 480), (135, 240), (100, 70)][trial % 3]
        img = _textured(h, w, rng, [2, 4, 8][trial // 2 % 3])
        if trial % 3 == 0:
            img = (img.astype(np.int32) + rng.integers(-40, 40, img.shape)).clip(0, 255).astype(np.uint8)
        ang = rng.uniform(-0.03, 0.03); tx, ty = rng.uniform(-6, 6, 2)
        M = np.float32([[np.cos(ang), np.sin(ang), tx], [-np.sin(ang), np.cos(ang), ty]])
        img2 = cv2.warpAffine(img, M, (w, h), borderMode=cv2.BORDER_REFLECT)
        g1 = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY); g2 = cv2.cvtColor(img2, cv2.COLOR_BGR2GRAY)
        pts = rng.uniform([-3, -3], [w + 2, h + 2], (16, 2)).astype(np.float32)  # some start outside the image
        if trial % 2 == 0:
            pts = np.trunc(pts)
        ref, st, _ = cv2.calcOpticalFlowPyrLK(g1, g2, pts, None, **LK)
        out, s = O.lk_track_restated(g1, g2, pts)
        assert np.array_equal(s, st[:, 0])
        ok = st[:, 0] == 1
        assert np.array_equal(out[ok].view(np.int32), ref[ok].view(np.int32))
        total += int(ok.sum())
    assert total > 60


def test_lk_lane_reduction_order_pinned_by_min_eigenvalue():
    """OPTFLOW_LK_GET_MIN_EIGENVALS exposes the structure-tensor sums; on white noise they round in float,
    which pins the (l0+l2)+(l1+l3) lane reduction."""
    rng = np.random.default_rng(7)
    win = 15
    g = rng.integers(0, 256, (100, 120), dtype=np.uint8)
    pts = rng.uniform([10, 10], [110, 90], (60, 2)).astype(np.float32)
    _, _, err = cv2.calcOpticalFlowPyrLK(g, g, pts, None, winSize=(win, win), maxLevel=0, criteria=(3, 10, 0.03),
                                         flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS)
    I = O._bordered(g, win); ix, iy = O.scharr_restated(g)
    Dm = np.zeros(I.shape + (2,), np.int16); Dm[win:-win, win:-win, 0] = ix; Dm[win:-win, win:-win, 1] = iy
    F = np.float32
    for k in range(len(pts)):
        p = pts[k] - F(7.0)
        ipx = int(np.floor(p[0])); ipy = int(np.floor(p[1]))
        w = O._weights(F(p[0] - F(ipx)), F(p[1] - F(ipy)))
        dx = O._sample(Dm[..., 0], ipx + win, ipy + win, win, w, 14); dy = O._sample(Dm[..., 1], ipx + win, ipy + win, win, w, 14)
        A11 = F(O._lane_sum(dx * dx) * O._FLT_SCALE); A12 = F(O._lane_sum(dx * dy) * O._FLT_SCALE); A22 = F(O._lane_sum(dy * dy) * O._FLT_SCALE)
        dd = F(A11 - A22)
        me = F(F(F(A22 + A11) - np.sqrt(F(F(dd * dd) + F(F(F(4.0) * A12) * A12)))) / F(2 * win * win))
        assert me == err[k, 0]


@needs_reference
def test_flow_filter_and_calibration_match_reference_methods():
    m = RH.bare_reference_model()
    W, H = 1280, 720
    c = S.make_flow_clip(5, W, H, seed=3, pan_px=4.0)
    fr = c["frames"]
    kp = SY.synthesize(D.decode_frame(c["heatmaps"][0], W, H, 0.3))
    prev_gray = cv2.cvtColor(fr[0], cv2.COLOR_BGR2GRAY)
    for i in range(1, 5):
        g = cv2.cvtColor(fr[i], cv2.COLOR_BGR2GRAY)
        ref = m.calculate_optical_flow(fr[i], prev_gray, kp, g)
        mine = O.calculate_optical_flow_restated(fr[i], O.gray_restated(fr[i - 1]), kp, O.gray_restated(fr[i]))
        assert list(ref.items()) == list(mine.items()) and len(ref) >= 10
        cal_ref = m.calibrate_keypoints(fr[i], ref)
        cal = O.calibrate_keypoints_restated(fr[i], mine)
        assert list(cal_ref.items()) == list(cal.items())
        assert [type(v[0]) for v in cal_ref.values()] == [type(v[0]) for v in cal.values()]
        kp = ref; prev_gray = g
    # the edge quirk: a dim keypoint in column 0 raises IndexError in both
    dark = np.zeros((40, 40, 3), np.uint8)
    with pytest.raises(IndexError):
        m.calibrate_keypoints(dark, {"X": (0, 20)})
    with pytest.raises(IndexError):
        O.calibrate_keypoints_restated(dark, {"X": (0, 20)})
    assert m.calibrate_keypoints(dark, {"X": (1, 1), "Y": (39, 39), "Z": (50, 3)}) == \
        O.calibrate_keypoints_restated(dark, {"X": (1, 1), "Y": (39, 39), "Z": (50, 3)})


@needs_reference
@pytest.mark.parametrize("fps,nh,nk,cal", [(8, 1, 2, False), (8, 1, 2, True), (6, 3, 1, False)])
def test_propagated_pipeline_json_identical_to_reference(fps, nh, nk, cal):
    W, H = 1280, 720
    c = S.make_flow_clip(14, W, H, seed=10 + fps + nk, pan_px=4.0)
    ref, rec = RH.run_reference(c["frames"], c["heatmaps"], c["objects"], fps=fps, num_homography=nh, num_keypoint_detection=nk,
                                calibration=cal)
    mine = P.get_coordinates_propagated(RH.stamp_frames(c["frames"]), c["heatmaps"], c["objects"], fps, nh, nk, calibration=cal)
    assert json.dumps(ref, default=float) == json.dumps(mine, default=float)
    assert min(len(ref[i]["Keypoints"]) for i in ref) >= 4


@pytest.mark.parametrize("cal", [False, True])
def test_oracle_reproduces_flow_golden(golden_dir, cal):
    """tests/golden/ref_flow_360p.npz holds what the unmodified reference returned (oracle/make_golden.py)."""
    import hashlib
    import os
    g = np.load(os.path.join(golden_dir, "ref_flow_360p.npz"))
    clip = S.make_flow_clip(int(g["n_frames"]), int(g["width"]), int(g["height"]), seed=int(g["seed"]), pan_px=float(g["pan_px"]))
    frames = RH.stamp_frames(clip["frames"])
    assert hashlib.sha256(np.stack(frames).tobytes()).hexdigest() == str(g["frames_sha256"])
    mine = P.get_coordinates_propagated(frames, clip["heatmaps"], clip["objects"], int(g["fps"]), int(g["num_homography"]),
                                        int(g["num_keypoint_detection"]), calibration=cal)
    assert json.dumps(mine, default=float, sort_keys=True) == str(g["result_json_cal" if cal else "result_json"])


def test_library_call_variant_equals_restated_pipeline():
    """oracle/pipeline.py with the cv2 calls themselves (the CPU baseline bench.py times) == the restated pieces."""
    c = S.make_flow_clip(12, 640, 360, seed=5, pan_px=2.0)
    c["heatmaps"][4] = 0.01
    a = P.get_coordinates_propagated(list(c["frames"]), c["heatmaps"], c["objects"], 8, 1, 2)
    b = P.get_coordinates_propagated(list(c["frames"]), c["heatmaps"], c["objects"], 8, 1, 2, library_calls=True)
    assert json.dumps(a, default=float) == json.dumps(b, default=float)
