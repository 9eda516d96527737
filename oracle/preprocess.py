"""Oracle: HRNet input preprocessing (test infrastructure).

``preprocess_reference_calls`` restates the reference's statements: cv2.cvtColor(BGR2RGB)
(eagle/models/coordinate_model.py:221) then ``A.Compose([A.Resize(540, 960), A.Normalize(),
ToTensorV2()])`` (:62-64, used :222 and :489-491).  albumentations 2.0.8 / albucore 0.0.24
(uv.lock:26-27,11-12) are absent from this image; their published behaviour for a uint8 HxWx3
image is: ``cv2.resize(img, (960, 540), interpolation=cv2.INTER_LINEAR)`` (uint8 out), then
``(img - 255*mean) * (1 / (255*std))`` in float32 with the ImageNet mean/std defaults, then
HWC -> CHW.

``resize_linear_u8_restated`` restates OpenCV's fixed-point INTER_LINEAR uint8 resize (imgproc
resize.cpp: 11-bit coefficients, HResizeLinear / VResizeLinear, and the INTER_LINEAR -> INTER_AREA
switch at exact 2x decimation) as integer numpy; pinned bit-exactly against the live cv2.resize in
tests/test_oracle_preprocess.py.  The CUDA kernel follows the same integer recipe.
"""
from __future__ import annotations

import cv2
import numpy as np

MODEL_H, MODEL_W = 540, 960  # A.Resize(540, 960), coordinate_model.py:63
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)  # A.Normalize() defaults
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def normalise_constants():
    """float32 (mean*255, 1/(std*255)) exactly as albucore forms them."""
    mean = (MEAN * np.float32(255.0)).astype(np.float32)
    denom = np.reciprocal(STD * np.float32(255.0), dtype=np.float32)
    return mean, denom


def preprocess_reference_calls(frame_bgr: np.ndarray) -> np.ndarray:
    """One BGR uint8 frame -> float32 (3, 540, 960), via the same library calls as the reference."""
    rgb = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2RGB)
    small = cv2.resize(rgb, (MODEL_W, MODEL_H), interpolation=cv2.INTER_LINEAR)
    mean, denom = normalise_constants()
    out = (small.astype(np.float32) - mean) * denom
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def _axis_taps(src: int, dst: int):
    """Per-output-index (offset, coef0, coef1) of OpenCV's linear resize along one axis."""
    scale = 1.0 / (dst / src)  # hal::resize: inv_scale = dsize/ssize, scale = 1./inv_scale
    ofs = np.empty(dst, np.int64)
    c0 = np.empty(dst, np.int64)
    c1 = np.empty(dst, np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        ofs[d] = s
        c0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        c1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return ofs, c0, c1


def resize_linear_u8_restated(img: np.ndarray, dst_w: int = MODEL_W, dst_h: int = MODEL_H) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC, restated."""
    H, W = img.shape[:2]
    if W == 2 * dst_w and H == 2 * dst_h:
        # exact 2x decimation is routed to the INTER_AREA fast path: rounded mean of each 2x2 block
        s = img.astype(np.int64)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xo, a0, a1 = _axis_taps(W, dst_w)
    yo, b0, b1 = _axis_taps(H, dst_h)
    # horizontal: left clamp zeroes the fraction, right clamp likewise (xmin/xmax handling)
    left = xo < 0
    xo = np.where(left, 0, xo); a0 = np.where(left, 2048, a0); a1 = np.where(left, 0, a1)
    right = xo >= W - 1
    xo = np.where(right, W - 1, xo); a0 = np.where(right, 2048, a0); a1 = np.where(right, 0, a1)
    x1 = np.minimum(xo + 1, W - 1)
    s = img.astype(np.int64)
    rows = s[:, xo] * a0[None, :, None] + s[:, x1] * a1[None, :, None]  # (H, dst_w, C) ints
    # vertical: row indices are clamped, coefficients are not
    y0 = np.clip(yo, 0, H - 1)
    y1 = np.clip(yo + 1, 0, H - 1)
    S0 = rows[y0]; S1 = rows[y1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess_restated(frame_bgr: np.ndarray) -> np.ndarray:
    """Same result as :func:`preprocess_reference_calls`, from the restated integer resize."""
    small = resize_linear_u8_restated(frame_bgr[:, :, ::-1])
    mean, denom = normalise_constants()
    out = (small.astype(np.float32) - mean) * denom
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def letterbox_geometry(h: int, w: int, imgsz: int = 960, stride: int = 32):
    """ultralytics 8.3.184 (uv.lock:1814-1815; third-party, absent here) ``LetterBox.__call__`` with the predictor's
    arguments (new_shape = imgsz, auto = True, scaleFill = False, scaleup = True, center = True), restated from its
    published source: (new_w, new_h, left, top, right, bottom).  PARITY UNPINNED against ultralytics itself."""
    r = min(imgsz / h, imgsz / w)
    new_unpad = int(round(w * r)), int(round(h * r))
    dw, dh = imgsz - new_unpad[0], imgsz - new_unpad[1]
    dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad[0], new_unpad[1], left, top, right, bottom


def letterbox_reference_calls(frame_bgr: np.ndarray, imgsz: int = 960, stride: int = 32) -> np.ndarray:
    """What ``self.detector_model(frame, ...)`` (eagle/models/coordinate_model.py:568) does to a BGR uint8 frame before the
    network runs: LetterBox (cv2.resize INTER_LINEAR if the size changes, cv2.copyMakeBorder with 114) and
    BasePredictor.preprocess (BGR -> RGB, HWC -> CHW, float32, / 255).  The cv2 calls are live; the sequence is the restated
    part.  Returns float32 (3, out_h, out_w)."""
    new_w, new_h, left, top, right, bottom = letterbox_geometry(frame_bgr.shape[0], frame_bgr.shape[1], imgsz, stride)
    img = frame_bgr
    if (img.shape[1], img.shape[0]) != (new_w, new_h):
        img = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    img = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
    chw = np.ascontiguousarray(img[..., ::-1].transpose(2, 0, 1))
    out = chw.astype(np.float32)
    out /= np.float32(255)
    return out
