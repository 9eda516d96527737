"""Mint the golden fixtures under tests/golden/ (run in the authoring container only).

    python -m oracle.make_golden

Every fixture is produced by EXECUTING the reference (oracle/ref_harness.py, /root/reference) or
the live third-party library the reference calls (cv2 4.13.0 here), never by the oracle restatement
or the CUDA path, so that both of those can be checked against it.  Inputs are regenerated from
seeds by eagle_b200.synthetic (too large to commit); each fixture stores a SHA-256 of the
regenerated input so that generator drift is detected instead of silently changing the test.
"""
from __future__ import annotations

import hashlib
import json
import os
import warnings

import cv2
import numpy as np

from eagle_b200 import synthetic
from eagle_b200.pitch import WORLD_XY_F32
from oracle import ref_harness

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cadence_fixture(name, n_frames, width, height, seed, fps, num_homography, blank):
    """Homography cadence (interval fps/num_homography, every frame a keypoint frame) incl. frames
    whose heatmaps are blanked so that the fit fails and the reference retries on the next frame."""
    clip = synthetic.make_clip(n_frames, width, height, seed=seed, with_frames=True, ghost_prob=0.05)
    synthetic.blank_heatmaps(clip["heatmaps"], blank)
    res, rec = ref_harness.run_reference(clip["frames"], clip["heatmaps"], clip["objects"], fps=fps,
                                         num_homography=num_homography, num_keypoint_detection=fps)
    np.savez_compressed(os.path.join(GOLDEN, name), n_frames=n_frames, width=width, height=height, seed=seed, fps=fps,
                        num_homography=num_homography, blank=np.array(blank, np.int32),
                        cv2_version=cv2.__version__, heatmaps_sha256=sha(clip["heatmaps"]),
                        result_json=json.dumps(res, default=float, sort_keys=True), n_fits=len(rec.fits))
    print(name, "frames", n_frames, "fits", len(rec.fits))


def clip_fixture(name, n_frames, width, height, seed, ghost_prob):
    clip = synthetic.make_clip(n_frames, width, height, seed=seed, with_frames=True, ghost_prob=ghost_prob)
    res, rec = ref_harness.run_reference(clip["frames"], clip["heatmaps"], clip["objects"])
    assert len(rec.fits) == n_frames, "fixture expects exactly one findHomography call per frame"
    fits_n = np.array([len(f["img_pts"]) for f in rec.fits], np.int32)
    nmax = int(fits_n.max())
    img = np.zeros((n_frames, nmax, 2), np.float32); wor = np.zeros((n_frames, nmax, 2), np.float32)
    msk = np.zeros((n_frames, nmax), np.uint8); Hs = np.full((n_frames, 3, 3), np.nan)
    for i, f in enumerate(rec.fits):
        n = fits_n[i]
        img[i, :n] = f["img_pts"]; wor[i, :n] = f["world_pts"]
        if f["H"] is not None:
            Hs[i] = f["H"]; msk[i, :n] = f["mask"].ravel()
    proj_pt = np.concatenate([p["pt"] for p in rec.projs]).astype(np.float32)
    proj_out = np.concatenate([p["out"] for p in rec.projs]).astype(np.float32)
    proj_H = np.stack([p["H"] for p in rec.projs])
    np.savez_compressed(
        os.path.join(GOLDEN, name),
        n_frames=n_frames, width=width, height=height, seed=seed, ghost_prob=ghost_prob,
        heatmaps_sha256=sha(clip["heatmaps"]), frames_sha256=sha(clip["frames"]),
        objects_json=json.dumps(clip["objects"], default=lambda o: o.tolist()),
        result_json=json.dumps(res, default=float, sort_keys=True),
        fit_n=fits_n, fit_img_pts=img, fit_world_pts=wor, fit_H=Hs, fit_mask=msk,
        proj_pt=proj_pt, proj_out=proj_out, proj_H=proj_H,
    )
    print(name, "frames", n_frames, "fits", len(rec.fits), "projs", len(rec.projs))


def decode_fixture():
    """Small heatmaps with the edge cases: ties (first max wins), maxima on the last row/column,
    scores at/around 0.01 and 0.3, duplicate pixel positions across channels.  Decoded by the
    reference's own get_keypoints / detect_keypoints code."""
    import importlib
    import torch
    cm = ref_harness.load_reference()
    hrnet = importlib.import_module("eagle.models.keypoint_hrnet")
    rng = np.random.default_rng(11)
    N, C, h, w = 6, 57, 12, 20
    hm = rng.uniform(0, 0.2, (N, C, h, w)).astype(np.float32)
    for n in range(N):
        for c in range(C):
            y, x = rng.integers(0, h), rng.integers(0, w)
            hm[n, c, y, x] = rng.choice([0.9, 0.3, 0.29999998, 0.30000001, 0.5, 1.0])
    hm[0, 3] = 0.004; hm[0, 4] = 0.01; hm[0, 5] = 0.0100001           # <= 0.01 dropped by get_keypoints
    hm[1, 6] = 0.7                                                      # all-equal map: argmax = 0
    hm[1, 7, h - 1, w - 1] = 0.95                                       # last element
    hm[1, 8, 2, 5] = hm[1, 8, 7, 1] = 0.99                              # tie -> first in row-major order
    hm[2, 10] = hm[2, 9]; hm[2, 11] = hm[2, 9]                          # three channels on one pixel, equal score
    hm[3, 12] = hm[3, 13] * np.float32(0.5) + np.float32(0.0)           # same argmax pixel, lower score
    hm[3, 14, :, :] = 0.1; hm[3, 14, 0, w - 1] = 0.8                    # x = w-1 -> xi == width
    hm[3, 15, :, :] = 0.1; hm[3, 15, h - 1, 0] = 0.8                    # y = h-1 -> yi == height

    class Fake(hrnet.KeypointModel):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.n_heatmaps = C

        def forward(self, x):
            return x

    kps = Fake().get_keypoints(torch.from_numpy(hm))
    flat = []
    for n, lst in enumerate(kps):
        for (i, xn, yn, sc) in lst:
            flat.append((n, i, xn, yn, sc))
    flat = np.array(flat, np.float64)
    # post-processing through the reference's detect_keypoints (coordinate_model.py:480-518)
    post = {}
    for (W, H) in [(1280, 720), (1920, 1080), (3840, 2160), (854, 480)]:
        model = cm.CoordinateModel.__new__(cm.CoordinateModel)
        model.keypoint_conf = 0.3

        class KM:
            unnormalized_model = [None, torch.nn.Conv2d(1, 1, 1)]

            def __init__(self, n):
                self.n = n

            def get_keypoints(self, frame):
                return [kps[self.n]]
        model.transforms = lambda image: {"image": torch.zeros(3, 4, 4)}
        for n in range(N):
            model.keypoint_model = KM(n)
            out = model.detect_keypoints(np.zeros((H, W, 3), np.uint8))
            post[f"{W}x{H}:{n}"] = {k: [int(v[0]), int(v[1])] for k, v in out.items()}
    np.savez_compressed(os.path.join(GOLDEN, "decode_small.npz"), heatmaps=hm, keypoints=flat,
                        postprocessed_json=json.dumps(post))
    print("decode_small", flat.shape, len(post))


def find_homography_fixture():
    """Point sets -> live cv2.findHomography(RANSAC, 5.0) results (the reference's call, :355)."""
    rng = np.random.default_rng(3)
    on = [i for i in range(57) if i not in (0, 1, 24, 25)]
    cases = []
    for t in range(96):
        W, H = [(1280, 720), (1920, 1080), (3840, 2160)][t % 3]
        cam = synthetic.sample_cameras(1, W, H, rng)[0]
        n = int(rng.integers(4, 54))
        sel = np.sort(rng.choice(on, n, replace=False))
        px = synthetic.project_points(cam, WORLD_XY_F32[sel].astype(np.float64))
        px += rng.normal(0, rng.choice([0.0, 0.5, 2.0]), px.shape)
        n_out = int(n * rng.uniform(0.0, 0.5)) if n > 6 else 0
        oi = rng.choice(n, n_out, replace=False)
        px[oi] = rng.uniform([0, 0], [W, H], (n_out, 2))
        img = np.rint(px).astype(np.float32)
        if t == 90:
            img[:] = np.c_[np.arange(n), 2 * np.arange(n)]  # all collinear -> None
        if t == 91:
            img[:] = img[0]  # all coincident -> None
        wor = WORLD_XY_F32[sel]
        Hm, mask = cv2.findHomography(img, wor, cv2.RANSAC, 5.0)
        cases.append((sel, img, Hm, mask))
    nmax = 53
    sel_a = np.full((len(cases), nmax), -1, np.int32); img_a = np.zeros((len(cases), nmax, 2), np.float32)
    H_a = np.full((len(cases), 3, 3), np.nan); m_a = np.zeros((len(cases), nmax), np.uint8); n_a = np.zeros(len(cases), np.int32)
    for i, (sel, img, Hm, mask) in enumerate(cases):
        n = len(sel); n_a[i] = n; sel_a[i, :n] = sel; img_a[i, :n] = img
        if Hm is not None:
            H_a[i] = Hm; m_a[i, :n] = mask.ravel()
    np.savez_compressed(os.path.join(GOLDEN, "find_homography_cv2.npz"), n=n_a, channels=sel_a, img_pts=img_a, H=H_a, mask=m_a,
                        cv2_version=cv2.__version__)
    print("find_homography_cv2", len(cases), "none:", int(np.isnan(H_a[:, 0, 0]).sum()))


def cascade_fixture(want=((-1, 120), (0, 40), (1, 100), (2, 100))):
    """The whole cascade of coordinate_model.py:354-357 on live cv2 for point sets on which the RANSAC leg returns None
    (families of tools/cascade_census.py: outlier-heavy camera views, points on one or two pitch lines, garbage,
    near-collinear / repeated / tiny-spread pixels) plus a few ordinary ones: which leg answered (want = how many sets
    per answering leg, -1 = none did), its H and mask."""
    from tools.cascade_census import FAMILIES, _world, make_set
    world, on = _world()
    rng = np.random.default_rng(2024)
    want, have = dict(want), {k: 0 for k, _ in want}
    cases, t = [], 0
    while any(have[k] < want[k] for k in want):
        fam = FAMILIES[t % len(FAMILIES)]; t += 1
        img, wor, sel = make_set(rng, fam, world, on, return_sel=True)
        if len(img) < 5:
            continue
        leg, Hm, mask = -1, None, None
        for k, (method, thr) in enumerate(((cv2.RANSAC, 5.0), (cv2.RHO, None), (cv2.LMEDS, None))):
            Hm, mask = cv2.findHomography(img, wor, method, thr)
            if Hm is not None:
                leg = k
                break
        if have[leg] >= want[leg]:
            continue
        have[leg] += 1
        cases.append((sel, img, leg, Hm, mask))
    nmax = 16
    T = len(cases)
    sel_a = np.full((T, nmax), -1, np.int32); img_a = np.zeros((T, nmax, 2), np.float32); n_a = np.zeros(T, np.int32)
    H_a = np.full((T, 3, 3), np.nan); m_a = np.zeros((T, nmax), np.uint8); leg_a = np.zeros(T, np.int32)
    for i, (sel, img, leg, Hm, mask) in enumerate(cases):
        n = len(sel); n_a[i] = n; sel_a[i, :n] = sel; img_a[i, :n] = img; leg_a[i] = leg
        if Hm is not None:
            H_a[i] = Hm; m_a[i, :n] = mask.ravel()
    np.savez_compressed(os.path.join(GOLDEN, "cascade_cv2.npz"), n=n_a, channels=sel_a, img_pts=img_a, leg=leg_a, H=H_a, mask=m_a,
                        cv2_version=cv2.__version__)
    print("cascade_cv2", T, "by leg (-1 none, 0 RANSAC, 1 RHO, 2 LMEDS):", {int(k): int((leg_a == k).sum()) for k in np.unique(leg_a)})


def find_hard_frames(width, height, want, seed=77):
    """Heatmap peak sets (channel -> heatmap pixel) whose decoded + synthesised keypoints make the RANSAC leg of
    coordinate_model.py:354-357 return None while RHO (leg 1) or LMEDS (leg 2) answers: `want` = {leg: how many}."""
    from oracle import decode as _decode, homography as _hom, synthesis as _syn
    from tools.cascade_census import _world
    world, on = _world()
    lines = {}
    for i in on:
        lines.setdefault(("x", round(float(world[i, 0]), 3)), []).append(i)
        lines.setdefault(("y", round(float(world[i, 1]), 3)), []).append(i)
    lines = [v for v in lines.values() if len(v) >= 3]
    rng = np.random.default_rng(seed)
    want = dict(want)
    found = []
    hh, ww = 135, 240
    while any(v > 0 for v in want.values()):
        fam = lines[int(rng.integers(len(lines)))]
        m = int(rng.integers(3, min(len(fam), 6) + 1))
        on_line = rng.choice(fam, m, replace=False)
        rest = [i for i in on if i not in fam]
        strays = rng.choice(rest, int(rng.integers(1, 4)), replace=False)
        a, b = rng.uniform([10, 10], [ww - 10, hh - 10]), rng.uniform([10, 10], [ww - 10, hh - 10])
        peaks = np.zeros((57, 3), np.int32)
        for c, t in zip(on_line, np.sort(rng.uniform(0, 1, m))):
            p = a + t * (b - a) + rng.normal(0, 0.4, 2)
            peaks[c] = (1, int(np.clip(round(p[1]), 0, hh - 1)), int(np.clip(round(p[0]), 0, ww - 1)))
        for c in strays:
            peaks[c] = (1, int(rng.integers(0, hh)), int(rng.integers(0, ww)))
        hm = np.zeros((57, hh, ww), np.float32)
        for c in range(57):
            if peaks[c, 0]:
                hm[c, peaks[c, 1], peaks[c, 2]] = 0.9
        kps = _decode.decode_frame(hm, width, height)
        if len(kps) >= 2:
            kps = _syn.synthesize(kps)
        img, wor, _ = _hom.gather_correspondences(kps)
        if len(img) < 5:
            continue
        leg = -1
        for k, (method, thr) in enumerate(((cv2.RANSAC, 5.0), (cv2.RHO, None), (cv2.LMEDS, None))):
            if cv2.findHomography(img, wor, method, thr)[0] is not None:
                leg = k
                break
        if want.get(leg, 0) > 0:
            want[leg] -= 1
            found.append((leg, peaks))
    return found


def apply_hard_frames(heatmaps, frames, peaks):
    for f, pk in zip(frames, peaks):
        heatmaps[f] = 0.0
        for c in range(57):
            if pk[c, 0]:
                heatmaps[f, c, pk[c, 1], pk[c, 2]] = 0.9


def cascade_clip_fixture(name="ref_cascade_clip_720p.npz", n_frames=14, width=1280, height=720, seed=21):
    """A clip on which the UNMODIFIED reference itself falls through to cv2.RHO and cv2.LMEDS (:354-357): four of its
    frames carry keypoint sets the RANSAC leg gives up on (two answered by RHO, two by LMEDS), every frame a keypoint
    and homography frame.  The dict it returns is what the CUDA path has to reproduce, rescued frames included."""
    hard = find_hard_frames(width, height, {1: 2, 2: 2})
    frames_idx = [2, 5, 8, 11]
    clip = synthetic.make_clip(n_frames, width, height, seed=seed, with_frames=True, ghost_prob=0.05)
    peaks = np.stack([pk for _, pk in hard])
    apply_hard_frames(clip["heatmaps"], frames_idx, peaks)
    res, rec = ref_harness.run_reference(clip["frames"], clip["heatmaps"], clip["objects"])
    methods = [f.get("method") for f in rec.fits]
    np.savez_compressed(os.path.join(GOLDEN, name), n_frames=n_frames, width=width, height=height, seed=seed,
                        hard_frames=np.array(frames_idx, np.int32), hard_peaks=peaks, hard_legs=np.array([l for l, _ in hard], np.int32),
                        cv2_version=cv2.__version__, heatmaps_sha256=sha(clip["heatmaps"]),
                        result_json=json.dumps(res, default=float, sort_keys=True), n_fit_calls=len(rec.fits))
    print(name, "frames", n_frames, "findHomography calls", len(rec.fits), "legs of the hard frames", [l for l, _ in hard], methods[:0])


def resize_fixture():
    """Checksums of the live cv2.resize(..., (960,540), INTER_LINEAR) output on seeded frames."""
    out = {}
    for (W, H) in [(1280, 720), (1920, 1080), (3840, 2160), (854, 480)]:
        fr = np.random.default_rng(W).integers(0, 256, (H, W, 3), dtype=np.uint8)
        small = cv2.resize(cv2.cvtColor(fr, cv2.COLOR_BGR2RGB), (960, 540), interpolation=cv2.INTER_LINEAR)
        out[f"{W}x{H}"] = {"seed": W, "frame_sha256": sha(fr), "resized_rgb_sha256": sha(small),
                           "head": small[:2, :8].reshape(-1).tolist()}
    with open(os.path.join(GOLDEN, "resize_cv2.json"), "w") as f:
        json.dump({"cv2_version": cv2.__version__, "cases": out}, f, indent=1)
    print("resize_cv2", list(out))


def flow_fixture(name, n_frames, width, height, seed, fps, num_homography, num_keypoint_detection, pan_px):
    """Sparse keypoint cadence (main.py:27 uses num_keypoint_detection=3): a rendered clip with a moving
    camera, so that the reference's Lucas-Kanade propagation, its filters and (second run) the brightness
    calibration all do real work.  The frames are reproducible from the seed; their hash is stored."""
    clip = synthetic.make_flow_clip(n_frames, width, height, seed=seed, pan_px=pan_px)
    out = {}
    for cal in (False, True):
        res, rec = ref_harness.run_reference(clip["frames"], clip["heatmaps"], clip["objects"], fps=fps, num_homography=num_homography,
                                             num_keypoint_detection=num_keypoint_detection, calibration=cal)
        out["result_json_cal" if cal else "result_json"] = json.dumps(res, default=float, sort_keys=True)
        print(name, "calibration", cal, "fits", len(rec.fits), "keypoints/frame", [len(res[i]["Keypoints"]) for i in res])
    stamped = np.stack(ref_harness.stamp_frames(clip["frames"]))
    np.savez_compressed(os.path.join(GOLDEN, name), n_frames=n_frames, width=width, height=height, seed=seed, fps=fps,
                        num_homography=num_homography, num_keypoint_detection=num_keypoint_detection, pan_px=pan_px,
                        frames_sha256=sha(stamped), heatmaps_sha256=sha(clip["heatmaps"]), **out)


def main():
    warnings.simplefilter("ignore")
    os.makedirs(GOLDEN, exist_ok=True)
    clip_fixture("ref_clip_720p.npz", 8, 1280, 720, seed=7, ghost_prob=0.05)
    clip_fixture("ref_clip_1080p.npz", 6, 1920, 1080, seed=8, ghost_prob=0.10)
    cadence_fixture("ref_cadence_720p.npz", 17, 1280, 720, seed=9, fps=5, num_homography=1, blank=[])
    # the reference itself goes through retry-after-failure: frame 0 and the scheduled frame 5 fail, so do the retries
    # on 6; 11 is a failed frame the cadence never asks for
    cadence_fixture("ref_cadence_retry_720p.npz", 17, 1280, 720, seed=9, fps=5, num_homography=1, blank=[0, 5, 6, 11])
    flow_fixture("ref_flow_360p.npz", 26, 640, 360, seed=31, fps=24, num_homography=1, num_keypoint_detection=3, pan_px=2.0)
    decode_fixture()
    find_homography_fixture()
    cascade_fixture()
    cascade_clip_fixture()
    resize_fixture()


if __name__ == "__main__":
    main()
