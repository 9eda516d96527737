"""Run the UNMODIFIED reference ``CoordinateModel.get_coordinates`` on synthetic inputs.

Test infrastructure, usable only where /root/reference is mounted (the authoring container; the
GPU box does not have it).  Used by oracle/make_golden.py to mint tests/golden/*.npz and by
tests/test_oracle_vs_reference.py to check the oracle restatement against the real thing.

How: the reference imports ultralytics / boxmot / albumentations at module import
time (coordinate_model.py:5,12-14), none of which exist offline, so empty stub modules are
registered first; the model object is created with ``__new__`` (skipping ``__init__`` and its
weight loading, :49-74) and given
  * a ``KeypointModel`` subclass whose ``forward`` returns pre-rendered heatmaps, so that the
    reference's OWN ``get_keypoints`` (keypoint_hrnet.py:575-595) decodes them,
  * ``transforms`` restating A.Resize(540,960)+A.Normalize()+ToTensorV2 with the same cv2 call,
  * ``detect_objects`` returning the synthetic boxes.
``cv2.findHomography`` / ``cv2.perspectiveTransform`` are wrapped with recorders because the
reference returns neither H, nor the mask, nor the un-truncated projections.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    import os
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "eagle", "models"))


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    mod("ultralytics", YOLO=_Dummy)
    mod("boxmot", BotSort=_Dummy)
    alb = mod("albumentations", Compose=_Dummy, Resize=_Dummy, Normalize=_Dummy)
    alb.pytorch = mod("albumentations.pytorch", ToTensorV2=_Dummy)


def load_reference():
    """Import and return the reference's coordinate_model module (cached)."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("eagle.models.coordinate_model")


class Recorder:
    """Wraps cv2.findHomography / cv2.perspectiveTransform (module attributes, so the reference's
    ``cv2.`` lookups see the wrappers) and logs every call."""

    def __init__(self):
        import cv2
        self.cv2 = cv2
        self.fits = []   # dict(img_pts, world_pts, method, H, mask)
        self.projs = []  # dict(pt, H, out)
        self._fh = cv2.findHomography
        self._pt = cv2.perspectiveTransform

    def __enter__(self):
        def fh(src, dst, method=0, thr=3.0, *a, **k):
            H, mask = self._fh(src, dst, method, thr, *a, **k)
            self.fits.append(dict(img_pts=np.array(src, copy=True), world_pts=np.array(dst, copy=True), method=method,
                                  H=None if H is None else H.copy(), mask=None if mask is None else mask.copy()))
            return H, mask

        def pt(src, H, *a, **k):
            out = self._pt(src, H, *a, **k)
            self.projs.append(dict(pt=np.array(src, copy=True).reshape(-1, 2), H=np.array(H, copy=True), out=out.copy().reshape(-1, 2)))
            return out

        self.cv2.findHomography = fh
        self.cv2.perspectiveTransform = pt
        return self

    def __exit__(self, *exc):
        self.cv2.findHomography = self._fh
        self.cv2.perspectiveTransform = self._pt
        return False


def stamp_frames(frames):
    """Copies of the frames with the frame index written into pixel (0,0) (see run_reference)."""
    stamped = []
    for i, fr in enumerate(frames):
        fr = np.array(fr, copy=True)
        fr[0, 0] = (i & 255, (i >> 8) & 255, (i >> 16) & 255)  # BGR
        stamped.append(fr)
    return stamped


def bare_reference_model(keypoint_conf: float = 0.3):
    """A reference CoordinateModel without networks: enough for calculate_optical_flow (:419-478),
    calibrate_keypoints (:520-555) and _synthesize_keypoints_with_line_intersections (:76-186)."""
    import cv2
    cm = load_reference()
    model = cm.CoordinateModel.__new__(cm.CoordinateModel)
    model.keypoint_conf = keypoint_conf
    model.detector_conf = 0.35
    model.lk_params = dict(winSize=(15, 15), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    return model


def run_reference(frames, heatmaps, objects_per_frame, fps: int = 1, num_homography: int = 1,
                  num_keypoint_detection: int = 1, keypoint_conf: float = 0.3, calibration: bool = False):
    """Call the real get_coordinates.  frames: sequence of (H,W,3) uint8; heatmaps (F,57,h,w) f32.

    With the default fps=1 both cadence intervals are 1 (coordinate_model.py:205-206): every frame
    is decoded and fitted and optical flow is never consulted (as long as >= 4 keypoints decode).
    Returns (result dict, Recorder).
    """
    import cv2
    import torch
    cm = load_reference()
    hrnet = importlib.import_module("eagle.models.keypoint_hrnet")

    hm_t = torch.from_numpy(np.ascontiguousarray(heatmaps))
    mean = np.array([0.485, 0.456, 0.406], np.float32) * 255.0
    denom = 1.0 / (np.array([0.229, 0.224, 0.225], np.float32) * 255.0)

    class FakeKeypointModel(hrnet.KeypointModel):
        """forward() returns the synthetic heatmaps in call order; everything else is the reference's."""

        def __init__(self):
            torch.nn.Module.__init__(self)
            self.n_heatmaps = 57
            self.unnormalized_model = torch.nn.Sequential(torch.nn.Identity(), torch.nn.Conv2d(1, 1, 1))

        def forward(self, x):
            # The reference evaluates ``mem.get(i, self.detect_keypoints(frame))`` eagerly
            # (coordinate_model.py:285), i.e. it runs the network a second time on every keypoint
            # frame, so forward() cannot hand out heatmaps in call order: the frame index rides in
            # element [0,0,0] of each input tensor instead (written by ``transforms`` below).
            idx = x[:, 0, 0, 0].round().long()
            return hm_t[idx]

    def transforms(image):
        small = cv2.resize(image, (960, 540), interpolation=cv2.INTER_LINEAR)
        t = (small.astype(np.float32) - mean) * denom
        t = np.ascontiguousarray(t.transpose(2, 0, 1))
        r, g, b = (int(v) for v in image[0, 0])  # RGB here (after the reference's cvtColor)
        t[0, 0, 0] = float((r << 16) | (g << 8) | b)  # frame index, stamped into pixel (0,0) below
        return {"image": torch.from_numpy(t)}

    model = cm.CoordinateModel.__new__(cm.CoordinateModel)
    model.keypoint_model = FakeKeypointModel()
    model.transforms = transforms
    model.keypoint_conf = keypoint_conf
    model.detector_conf = 0.35
    model.lk_params = dict(winSize=(15, 15), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    it = iter(objects_per_frame)
    model.detect_objects = lambda frame: next(it)

    stamped = stamp_frames(frames)
    with Recorder() as rec:
        res = model.get_coordinates(stamped, fps=fps, num_homography=num_homography,
                                    num_keypoint_detection=num_keypoint_detection, verbose=False, calibration=calibration)
    return res, rec
