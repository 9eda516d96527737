"""GPU parity tests of the keypoint-propagation kernels (csrc/flow.cu) and of the sparse keypoint cadence
end to end, against oracle/optflow.py + oracle/pipeline.py (pinned to cv2 / numpy / the reference on the
CPU side by tests/test_oracle_optflow.py) and against live cv2 where the box has it.  Everything here is
integer or bit-pattern equality."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from eagle_b200 import synthetic  # noqa: E402
from oracle import optflow as O  # noqa: E402
from oracle import pipeline  # noqa: E402


@pytest.fixture(scope="module")
def engine():
    from eagle_b200.engine import GeometryEngine
    return GeometryEngine("cuda:0")


def _layout(H, W, max_level=2, win=15):
    lv = [(H, W)]
    for _ in range(max_level):
        h, w = (lv[-1][0] + 1) // 2, (lv[-1][1] + 1) // 2
        if w <= win or h <= win:
            break
        lv.append((h, w))
    offs, o = [], 0
    for h, w in lv:
        offs.append(o)
        o += (h * w + 15) // 16 * 16
    return lv, offs, o


def _textured(h, w, rng, cell):
    import cv2
    base = rng.integers(0, 256, (h // cell + 2, w // cell + 2, 3), dtype=np.uint8)
    return cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)


def _keypoint_set(points_per_frame, dev="cuda"):
    """points_per_frame: list of (channel order list, {channel: (x, y)})."""
    from eagle_b200.engine import KeypointSet
    n = len(points_per_frame)
    xy = np.zeros((n, 57, 2), np.int32); order = np.zeros((n, 64), np.uint8); count = np.zeros((n, 2), np.int32)
    for f, (chs, pts) in enumerate(points_per_frame):
        for j, c in enumerate(chs):
            order[f, j] = c
            xy[f, c] = pts[c]
        count[f] = (len(chs), len(chs))
    t = lambda a: torch.from_numpy(a).to(dev)
    return KeypointSet(None, None, t(xy), t(order), t(count), torch.zeros((n, 64), dtype=torch.uint8, device=dev))



def test_strided_pyramid_equals_dense(engine):
    """egl_gray_pyramid_strided: the pyramids of one step of every chain at a time (frames s, s + k, ...) fill the same
    buffer, byte for byte, as one dense pass -- at an interval that does not divide the clip and with a leading frame."""
    rng = np.random.default_rng(4)
    F, H, W, k = 23, 96, 160, 4
    frames = torch.from_numpy(rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)).cuda()
    dense = engine.gray_pyramid(frames, 2)
    for base in (0, 1):
        pyr = engine.alloc_pyramid(F, H, W, 2)
        pyr.fill_(0xAB)
        if base:
            engine.gray_pyramid_step(frames, pyr, 0, F, 2)
        for s in range(k):
            engine.gray_pyramid_step(frames, pyr, base + s, k, 2)
        torch.cuda.synchronize()
        assert torch.equal(pyr, dense), base

@pytest.mark.parametrize("shape", [(77, 101), (360, 640), (1080, 1920), (33, 18), (90, 160), (101, 256), (37, 48), (64, 1280), (720, 1280)])
def test_gray_pyramid_matches_oracle(engine, shape):
    H, W = shape
    rng = np.random.default_rng(H)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    pyr = engine.gray_pyramid(torch.from_numpy(frames).cuda()).cpu().numpy()
    lv, offs, nbytes = _layout(H, W)
    assert pyr.shape == (3, nbytes)
    for f in range(3):
        want = O.build_pyramid(O.gray_restated(frames[f]), 15, 2)
        assert len(want) == len(lv)
        for (h, w), o, img in zip(lv, offs, want):
            assert np.array_equal(pyr[f, o:o + h * w].reshape(h, w), img), (f, h, w)


def test_tracker_bit_exact_vs_cv2_and_oracle(engine):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    for (H, W) in [(270, 480), (100, 70), (360, 640)]:
        pairs = 5
        frames = np.empty((2 * pairs, H, W, 3), np.uint8)
        sets = []
        for p in range(pairs):
            img = _textured(H, W, rng, [2, 4, 8, 16, 3][p])
            if p % 2 == 0:
                img = (img.astype(np.int32) + rng.integers(-40, 40, img.shape)).clip(0, 255).astype(np.uint8)
            if p == 4:
                img = rng.integers(0, 256, img.shape, dtype=np.uint8)  # white noise: every float sum rounds
            ang = rng.uniform(-0.03, 0.03); tx, ty = rng.uniform(-6, 6, 2)
            M = np.float32([[np.cos(ang), np.sin(ang), tx], [-np.sin(ang), np.cos(ang), ty]])
            frames[2 * p] = img
            frames[2 * p + 1] = cv2.warpAffine(img, M, (W, H), borderMode=cv2.BORDER_REFLECT)
            npts = [57, 40, 12, 1, 57][p]
            chs = list(rng.permutation(57)[:npts])
            pts = {int(c): (int(rng.integers(-3, W + 3)), int(rng.integers(-3, H + 3))) for c in chs}  # some outside the image
            sets.append(([int(c) for c in chs], pts))
        dev_frames = torch.from_numpy(frames).cuda()
        pyr = engine.gray_pyramid(dev_frames)
        kp = _keypoint_set(sets)
        new_pts, status = engine.track(pyr, H, W, kp, 0, 1, 2)
        new_pts = new_pts.cpu().numpy(); status = status.cpu().numpy()
        tracked = 0
        for p, (chs, pts) in enumerate(sets):
            g1 = O.gray_restated(frames[2 * p]); g2 = O.gray_restated(frames[2 * p + 1])
            src = np.array([pts[c] for c in chs], np.float32)
            ref, st, _ = cv2.calcOpticalFlowPyrLK(g1, g2, src, None, **lk)
            assert np.array_equal(status[p, chs], st[:, 0]), (H, W, p)   # results are stored by channel
            ok = st[:, 0] == 1
            assert np.array_equal(new_pts[p, chs][ok].view(np.int32), ref[ok].view(np.int32)), (H, W, p)
            tracked += int(ok.sum())
            if p < 2:  # and the restated oracle (slow: two pairs per size)
                out, s = O.lk_track_restated(g1, g2, src)
                assert np.array_equal(s, st[:, 0]) and np.array_equal(out[ok].view(np.int32), ref[ok].view(np.int32))
        assert tracked > 60


def test_tracker_warp_and_thread_kernels_agree():
    """EGL_TRACK_VARIANT=1 runs lk_track_point (csrc/flow_core.cuh, the scalar statement the CPU suite compiles for
    the host and compares with live cv2) one thread per point; the default warp kernel has to produce the same
    bits.  Subprocesses, because the switch is read once per process.  The thread kernel only exists in builds with
    -DEGL_BENCH_VARIANTS (EGL_BENCH_VARIANTS=1 python -m eagle_b200.build --force; tools/variants_check.sh runs it)."""
    from eagle_b200 import _native
    if not (_native.lib.egl_build_flags() & 1):
        pytest.skip("libeagle_b200.so was built without -DEGL_BENCH_VARIANTS: the one-thread-per-point tracker is not in it")
    import subprocess
    import sys
    import tempfile
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from eagle_b200 import synthetic
from eagle_b200.engine import GeometryEngine, KeypointSet
e = GeometryEngine("cuda:0")
rng = np.random.default_rng(11)
outs = []
for (W, H) in [(640, 360), (202, 117)]:
    clip = synthetic.make_flow_clip(6, W, H, seed=9, pan_px=3.0)
    fr = clip["frames"].copy(); fr[4] = rng.integers(0, 256, fr[4].shape, dtype=np.uint8)   # one white-noise frame
    pyr = e.gray_pyramid(torch.from_numpy(fr).cuda())
    xy = np.zeros((5, 57, 2), np.int32); order = np.zeros((5, 64), np.uint8); count = np.zeros((5, 2), np.int32)
    for f in range(5):
        chs = rng.permutation(57)[:rng.integers(1, 58)]
        order[f, :len(chs)] = chs; count[f] = len(chs)
        xy[f, chs] = np.c_[rng.integers(-3, W + 3, len(chs)), rng.integers(-3, H + 3, len(chs))]
    t = lambda a: torch.from_numpy(a).cuda()
    kp = KeypointSet(None, None, t(xy), t(order), t(count), None)
    pts = torch.zeros((5, 64, 2), device="cuda"); st = torch.zeros((5, 64), dtype=torch.uint8, device="cuda")
    e.track(pyr, H, W, kp, 0, 1, 1, out=(pts, st))
    outs += [pts.cpu().numpy().ravel().view(np.int32).astype(np.int64), st.cpu().numpy().ravel().astype(np.int64)]
np.save(sys.argv[1], np.concatenate(outs))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for variant in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as tf:
            env = dict(os.environ, EGL_TRACK_VARIANT=variant)
            subprocess.run([sys.executable, "-c", code, tf.name], check=True, env=env, timeout=300)
            res.append(np.load(tf.name))
    assert np.array_equal(res[0], res[1]) and int((res[0] != 0).sum()) > 200


def test_filter_flow_matches_oracle(engine):
    from eagle_b200.engine import KeypointSet
    from eagle_b200.pitch import LANDMARK_NAMES
    W, H = 640, 360
    clip = synthetic.make_flow_clip(6, W, H, seed=4, pan_px=3.0)
    frames = clip["frames"]
    rng = np.random.default_rng(8)
    dev_frames = torch.from_numpy(frames).cuda()
    pyr = engine.gray_pyramid(dev_frames)
    sets = []
    for f in range(5):
        px, vis = synthetic.landmark_pixels(clip["cameras"][f], W, H)
        chs = [int(c) for c in rng.permutation(np.nonzero(vis)[0])]
        pts = {c: (int(px[c, 0]), int(px[c, 1])) for c in chs}
        if f == 1:  # a few far-off / hue-changing points so that every filter fires
            for c in chs[:3]:
                pts[c] = (int(rng.integers(0, W)), int(rng.integers(0, H // 4)))
        if f == 3:
            chs = chs[:2]
        if f == 4:
            chs = []
        sets.append((chs, pts))
    kp = _keypoint_set(sets)
    new_pts, status = engine.track(pyr, H, W, kp, 0, 1, 1)
    out = KeypointSet(None, None, torch.zeros((5, 57, 2), dtype=torch.int32, device="cuda"), torch.zeros((5, 64), dtype=torch.uint8, device="cuda"),
                      torch.zeros((5, 2), dtype=torch.int32, device="cuda"), torch.full((5, 64), 9, dtype=torch.uint8, device="cuda"))
    engine.filter_flow(dev_frames, 1, 1, kp, new_pts, status, out)
    xy = out.xy.cpu().numpy(); order = out.order.cpu().numpy(); count = out.count.cpu().numpy(); src = out.src.cpu().numpy()
    dropped = 0
    for f, (chs, pts) in enumerate(sets):
        prev = {LANDMARK_NAMES[c]: pts[c] for c in chs}
        want = O.calculate_optical_flow_restated(frames[f + 1], O.gray_restated(frames[f]), prev, O.gray_restated(frames[f + 1]))
        got = {LANDMARK_NAMES[c]: (int(xy[f, c, 0]), int(xy[f, c, 1])) for c in order[f, :count[f, 0]]}
        assert list(got.items()) == [(k, (int(v[0]), int(v[1]))) for k, v in want.items()], f
        assert all(src[f, c] == 1 for c in order[f, :count[f, 0]]) and int(src[f].sum()) == count[f, 0]
        dropped += len(chs) - len(want)
    assert dropped > 0


def test_merge_and_calibrate_match_oracle(engine):
    from eagle_b200.pitch import LANDMARK_NAMES
    rng = np.random.default_rng(3)
    # merge == {**a, **b}
    A, B = [], []
    for f in range(40):
        ca = [int(c) for c in rng.permutation(57)[:rng.integers(0, 30)]]
        cb = [int(c) for c in rng.permutation(57)[:rng.integers(0, 30)]]
        A.append((ca, {c: (int(rng.integers(0, 999)), int(rng.integers(0, 999))) for c in ca}))
        B.append((cb, {c: (int(rng.integers(0, 999)), int(rng.integers(0, 999))) for c in cb}))
    a = _keypoint_set(A); b = _keypoint_set(B); b.src.fill_(1)
    apply = torch.ones(40, dtype=torch.uint8, device="cuda"); apply[7] = 0
    engine.merge(a, b, apply)
    xy = a.xy.cpu().numpy(); order = a.order.cpu().numpy(); count = a.count.cpu().numpy(); src = a.src.cpu().numpy()
    for f in range(40):
        da = {c: A[f][1][c] for c in A[f][0]}; db = {c: B[f][1][c] for c in B[f][0]}
        want = {**da, **db} if f != 7 else da
        got = {int(c): (int(xy[f, c, 0]), int(xy[f, c, 1])) for c in order[f, :count[f, 0]]}
        assert list(got.items()) == list(want.items())
        assert all(src[f, c] == (1 if (c in db and f != 7) else 0) for c in got)
    # calibration
    W, H = 320, 200
    frames = rng.integers(0, 256, (4, H, W, 3), dtype=np.uint8)
    frames[1] //= 3   # mostly dim: most keypoints move
    frames[2, :3] = 0; frames[2, :, :3] = 0
    sets = []
    for f in range(4):
        chs = [int(c) for c in rng.permutation(57)[:45]]
        pts = {c: (int(rng.integers(1, W + 4)), int(rng.integers(1, H + 4))) for c in chs}
        for c in chs[:6]:  # edge cases: near the right / bottom borders, rows/cols 1..3
            pts[c] = (int(rng.choice([1, 2, 3, W - 1, W - 2])), int(rng.choice([1, 2, 3, H - 1, H - 2])))
        sets.append((chs, pts))
    sets[3][1][sets[3][0][0]] = (0, 50)  # dim pixel in column 0 -> IndexError in the reference
    frames[3, 50, 0] = 10
    kp = _keypoint_set(sets)
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    engine.calibrate(torch.from_numpy(frames).cuda(), 0, 1, kp, err)
    xy = kp.xy.cpu().numpy(); src = kp.src.cpu().numpy()
    assert err.cpu().tolist() == [0, 0, 0, 1]
    moved = 0
    for f in range(3):
        chs, pts = sets[f]
        want = O.calibrate_keypoints_restated(frames[f], {LANDMARK_NAMES[c]: pts[c] for c in chs})
        for c in chs:
            v = want[LANDMARK_NAMES[c]]
            assert (int(v[0]), int(v[1])) == (int(xy[f, c, 0]), int(xy[f, c, 1])), (f, c)
            assert src[f, c] == (1 if isinstance(v[0], np.integer) else 0)
            moved += isinstance(v[0], np.integer)
    assert moved > 20
    with pytest.raises(IndexError):
        O.calibrate_keypoints_restated(frames[3], {LANDMARK_NAMES[c]: sets[3][1][c] for c in sets[3][0]})


class _Net:
    """Stand-in for the keypoint network: recognises which frame a preprocessed tensor came from and
    returns that frame's pre-rendered heatmaps (the path may ask for any frame, in any order)."""

    def __init__(self, engine, frames, heatmaps):
        self.hm = torch.from_numpy(heatmaps).cuda()
        x = engine.preprocess(torch.from_numpy(np.ascontiguousarray(frames)).cuda())
        self.sig = self._signature(x)
        assert len(torch.unique(self.sig)) == len(frames)
        self.calls = []

    @staticmethod
    def _signature(x):
        return x[:, :, ::37, ::41].double().sum(dim=(1, 2, 3))

    def __call__(self, x):
        idx = [int(torch.nonzero(self.sig == s)[0, 0]) for s in self._signature(x)]
        self.calls.extend(idx)
        return self.hm[idx]


def _run_both(engine, clip, fps, nh, nk, cal=False, piece=None):
    from eagle_b200.coordinate_model import CoordinateModel
    frames = list(clip["frames"])
    net = _Net(engine, clip["frames"], clip["heatmaps"])
    objs = iter(clip["objects"])
    model = CoordinateModel(keypoint_model=net, detect_objects=lambda fr: next(objs), chunk=8)
    model.always_propagate = True
    if piece is not None:
        model.piece_frames = piece
    got = model.get_coordinates(frames, fps=fps, num_homography=nh, num_keypoint_detection=nk, verbose=False, calibration=cal)
    want = pipeline.get_coordinates_propagated(frames, clip["heatmaps"], clip["objects"], fps, nh, nk, calibration=cal)
    return got, want, model.last_stats, net


def _same(got, want):
    a = json.dumps(got, default=float); b = json.dumps(want, default=float)
    if a != b:
        for i in want:
            if json.dumps(got[i], default=float) != json.dumps(want[i], default=float):
                raise AssertionError(f"first difference at frame {i}:\n got  {got[i]['Keypoints']}\n want {want[i]['Keypoints']}")
    # json.dumps(default=float) equality covers the value types; also compare the Python types of the keypoint values
    for i in want:
        assert [type(v).__name__ + type(v[0]).__name__ for v in got[i]["Keypoints"].values()] == \
               [type(v).__name__ + type(v[0]).__name__ for v in want[i]["Keypoints"].values()], i


W, H = 640, 360


@pytest.mark.parametrize("fps,nh,nk,cal,n", [(8, 1, 2, False, 22), (8, 1, 2, True, 14), (6, 3, 1, False, 15), (24, 1, 3, False, 26), (5, 1, 5, False, 9)])
def test_sparse_cadence_clip_identical_to_oracle(engine, fps, nh, nk, cal, n):
    clip = synthetic.make_flow_clip(n, W, H, seed=100 + fps + nk, pan_px=2.0)
    got, want, stats, net = _run_both(engine, clip, fps, nh, nk, cal)
    _same(got, want)
    assert min(len(want[i]["Keypoints"]) for i in want) >= 4
    assert stats["repaired_chains"] == 0 and stats["fallback_frames"] == 0
    k = max(1, int(fps / max(1, nk)))
    assert sorted(set(net.calls)) == list(range(0, n, k))  # the network ran on the chain heads only


@pytest.mark.parametrize("n", [1, 2, 5])
def test_sparse_cadence_very_short_clips(engine, n):
    """Clips shorter than one chain (and one frame longer than a chain)."""
    clip = synthetic.make_flow_clip(n, W, H, seed=40 + n, pan_px=2.0)
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2)   # keypoint interval 4
    _same(got, want)
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2, cal=True, piece=4)
    _same(got, want)


def test_sparse_cadence_rescues_identical_to_oracle(engine):
    """The rare branches: first-frame rescue (:288-307), a head with < 4 landmarks (:308-311), flow that
    loses the scene (:316-320), a failed fit whose retry flag crosses a chain boundary (:333)."""
    # first frames blank -> backward flow from the first good frame
    clip = synthetic.make_flow_clip(14, W, H, seed=21, pan_px=2.0); clip["heatmaps"][0:3] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2)
    _same(got, want); assert stats["first_frame_rescue"] and stats["repaired_chains"] >= 1
    # a blank head in the middle, and a head with only 3 landmarks
    clip = synthetic.make_flow_clip(14, W, H, seed=22, pan_px=2.0); clip["heatmaps"][4] = 0.01; clip["heatmaps"][8, 3:] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2)
    _same(got, want); assert stats["repaired_chains"] >= 2
    # scene cut to a black frame: flow keeps nothing -> the network is asked for that frame
    clip = synthetic.make_flow_clip(14, W, H, seed=23, pan_px=2.0)
    clip["frames"][6] = 0
    got, want, stats, net = _run_both(engine, clip, 8, 1, 2)
    _same(got, want); assert stats["fallback_frames"] >= 1 and any(c % 4 for c in net.calls)
    # nothing anywhere
    clip = synthetic.make_flow_clip(6, W, H, seed=24); clip["heatmaps"][:] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2)
    _same(got, want)
    # every frame a network frame, some of them blank (optical-flow rescue at interval 1)
    clip = synthetic.make_flow_clip(8, W, H, seed=25, pan_px=2.0); clip["heatmaps"][3] = 0.01; clip["heatmaps"][0] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 1, 1, 1)
    _same(got, want)
    # retry flag: the scheduled fit of frame 8 fails (3 landmarks, flow adds none on a black previous frame)
    clip = synthetic.make_flow_clip(14, W, H, seed=26, pan_px=2.0)
    clip["heatmaps"][8, 3:] = 0.01; clip["frames"][7] = 0; clip["heatmaps"][7] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2)
    _same(got, want)


def test_sparse_cadence_in_pieces_identical_to_oracle(engine):
    """Long clips are processed in pieces of whole chains with the boundary state carried over; with pieces of
    one or two chains every boundary effect lands on a piece boundary somewhere."""
    for piece in (4, 8):
        clip = synthetic.make_flow_clip(22, W, H, seed=110, pan_px=2.0)
        got, want, stats, _ = _run_both(engine, clip, 8, 1, 2, piece=piece)
        _same(got, want); assert stats["pieces"] == (22 + piece - 1) // piece
        # blank head + 3-landmark head: flow from the carried frame / keypoints joins in
        clip = synthetic.make_flow_clip(14, W, H, seed=22, pan_px=2.0); clip["heatmaps"][4] = 0.01; clip["heatmaps"][8, 3:] = 0.01
        got, want, stats, _ = _run_both(engine, clip, 8, 1, 2, piece=piece)
        _same(got, want); assert stats["repaired_chains"] >= 2
        # retry flag carried across a piece boundary; black frame -> fallback detection inside a later piece
        clip = synthetic.make_flow_clip(14, W, H, seed=26, pan_px=2.0)
        clip["heatmaps"][8, 3:] = 0.01; clip["frames"][7] = 0; clip["heatmaps"][7] = 0.01
        got, want, stats, _ = _run_both(engine, clip, 8, 1, 2, piece=piece)
        _same(got, want)
        # first-frame rescue inside the first piece, calibration on
        clip = synthetic.make_flow_clip(14, W, H, seed=21, pan_px=2.0); clip["heatmaps"][0:2] = 0.01
        got, want, stats, _ = _run_both(engine, clip, 8, 1, 2, cal=True, piece=piece)
        _same(got, want); assert stats["first_frame_rescue"]
    # the scan for the first usable frame runs past the first piece (frames 0-5 blank, pieces of 4): the adapter retries longer
    clip = synthetic.make_flow_clip(14, W, H, seed=27, pan_px=2.0); clip["heatmaps"][0:6] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 8, 1, 2, piece=4)
    _same(got, want); assert stats["first_frame_rescue"]
    # interval 1 in pieces of one frame: every frame is a piece
    clip = synthetic.make_flow_clip(8, W, H, seed=25, pan_px=2.0); clip["heatmaps"][3] = 0.01
    got, want, stats, _ = _run_both(engine, clip, 1, 1, 1, piece=1)
    _same(got, want); assert stats["pieces"] == 8
    clip["heatmaps"][0:2] = 0.01   # ... and the clip starts blank: the one-frame first piece has to grow
    got, want, stats, _ = _run_both(engine, clip, 1, 1, 1, piece=1)
    _same(got, want); assert stats["first_frame_rescue"]


def test_sparse_cadence_golden_from_reference(engine, golden_dir):
    """The dict the UNMODIFIED reference produced for a rendered clip at main.py's cadence
    (num_homography=1, num_keypoint_detection=3; oracle/make_golden.py::golden_flow_clip)."""
    from eagle_b200.coordinate_model import CoordinateModel
    from oracle.ref_harness import stamp_frames
    g = np.load(os.path.join(golden_dir, "ref_flow_360p.npz"))
    n, w, h, fps = int(g["n_frames"]), int(g["width"]), int(g["height"]), int(g["fps"])
    clip = synthetic.make_flow_clip(n, w, h, seed=int(g["seed"]), pan_px=float(g["pan_px"]))
    frames = stamp_frames(clip["frames"])
    import hashlib
    assert hashlib.sha256(np.stack(frames).tobytes()).hexdigest() == str(g["frames_sha256"])
    for cal in (False, True):
        net = _Net(engine, np.stack(frames), clip["heatmaps"])
        objs = iter(clip["objects"])
        model = CoordinateModel(keypoint_model=net, detect_objects=lambda fr: next(objs))
        got = model.get_coordinates(frames, fps=fps, num_homography=int(g["num_homography"]),
                                    num_keypoint_detection=int(g["num_keypoint_detection"]), verbose=False, calibration=cal)
        assert json.dumps(got, default=float, sort_keys=True) == str(g["result_json_cal" if cal else "result_json"])


def test_propagated_path_full_size_properties(engine):
    """2250-frame 1080p-sized state at main.py's cadence is too slow for the Python oracle; check the
    structural properties instead on a 96-frame 1080p clip: heads equal the per-frame path, every frame
    has a homography, flowed keypoints stay close to the true landmark positions."""
    from eagle_b200.coordinate_model import CoordinateModel
    from eagle_b200.pitch import LANDMARK_INDEX
    n, Wf, Hf = 48, 1920, 1080
    clip = synthetic.make_flow_clip(n, Wf, Hf, seed=7, pan_px=3.0)
    net = _Net(engine, clip["frames"], clip["heatmaps"])
    objs = iter(clip["objects"])
    model = CoordinateModel(keypoint_model=net, detect_objects=lambda fr: next(objs))
    got = model.get_coordinates(list(clip["frames"]), fps=24, num_homography=1, num_keypoint_detection=3, verbose=False)
    assert sorted(got) == list(range(n))
    worst = 0.0
    for i in range(n):
        assert got[i]["Boundaries"] != [None] * 4 and len(got[i]["Keypoints"]) >= 4
        px, _ = synthetic.landmark_pixels(clip["cameras"][i], Wf, Hf)
        for name, v in got[i]["Keypoints"].items():
            worst = max(worst, float(np.hypot(v[0] - px[LANDMARK_INDEX[name], 0], v[1] - px[LANDMARK_INDEX[name], 1])))
    assert worst < 40.0, worst


def _sharded_flow_worker(rank, world, port, scenario, q):
    import torch.distributed as dist
    from eagle_b200.engine import GeometryEngine
    from eagle_b200.sharding import chain_range, run_sharded_propagated
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = GeometryEngine("cuda:0")
        n, k, h, fps = (5 if scenario == "short" else 14), 4, 8, 8
        if scenario in ("plain", "short"):
            clip = synthetic.make_flow_clip(n, W, H, seed=110, pan_px=2.0)
        else:  # blank head + 3-landmark head on shard boundaries, retry flag crossing, fallback frame
            clip = synthetic.make_flow_clip(n, W, H, seed=22, pan_px=2.0)
            clip["heatmaps"][4] = 0.01; clip["heatmaps"][8, 3:] = 0.01; clip["frames"][10] = 0
        lo, hi = chain_range(n, k, rank, world)
        halo = 1 if lo > 0 else 0
        dev_frames = torch.from_numpy(np.ascontiguousarray(clip["frames"][lo - halo:hi])).cuda()
        net = _Net(eng, clip["frames"], clip["heatmaps"])
        if hi > lo:
            heads = net(eng.preprocess(dev_frames[halo::k].contiguous())).contiguous()
        else:   # a rank past the end of a short clip holds only its predecessor's last frame
            heads = torch.zeros((0, 57, 135, 240), device="cuda")
        detect = lambda i: net(eng.preprocess(dev_frames[halo + i - lo:halo + i - lo + 1].contiguous())).contiguous()
        res = run_sharded_propagated(eng, dev_frames, heads, detect, clip["objects"][lo:hi], fps, k, h, first_frame=lo)
        if rank == 0:
            want = pipeline.get_coordinates_propagated(list(clip["frames"]), clip["heatmaps"], clip["objects"], fps, 1, 2)
            q.put(json.dumps(res, default=float) == json.dumps(want, default=float))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,scenario", [(2, "plain"), (3, "rescues"), (3, "short")])
def test_sparse_cadence_sharded_between_chains(world, scenario):
    """Chains shard over ranks (here the ranks share cuda:0 and talk over gloo; NCCL on a multi-GPU box): the
    parallel passes run at once, the boundary state is handed from rank to rank, rank 0 assembles the dict."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_flow_worker, args=(r, world, port, scenario, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
