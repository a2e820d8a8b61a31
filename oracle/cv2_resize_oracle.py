"""CPU oracle for the image resampling steps on either side of the hot path.  TEST INFRASTRUCTURE ONLY.

The reference resizes with OpenCV, a third-party dependency that is not part of /root/reference
(`opencv-python==4.1.1.26`, /root/reference/requirements.txt:1):
  * Testing/dataloader.py:63   `cv2.resize(img, self.size)` on the uint8 RGB frame (INTER_LINEAR, the default)
  * Testing/test.py:64         `cv2.resize(pred, (W//4, H//4), interpolation=cv2.INTER_NEAREST)` on the int8 label map
This file restates OpenCV's published algorithm for those two calls (modules/imgproc/src/resize.cpp: the fixed-point
`INTER_LINEAR` path for 8-bit images with INTER_RESIZE_COEF_BITS = 11, and `resizeNN`) in numpy.

Pinning: tests/test_ingest.py checks this restatement bit for bit (a) against the cv2 that is installed in the build /
GPU image (4.13.0 -- the 8-bit algorithm has not changed since the reference's 4.1.1) on random images over up- and
down-scaling shapes including 1024x2048 -> 769x1537, and (b) against tests/golden/resize_cases.npz, outputs of cv2
itself stored by tests/golden/make_golden_resize.py.  Parity status: pinned against outputs of the dependency itself.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11                      # INTER_RESIZE_COEF_BITS
COEF_SCALE = 1 << COEF_BITS         # 2048


def _positions(src: int, dst: int):
    """fx = (float)((dx + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in double; sx = cvFloor(fx); fx -= sx."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    return s, (f - s.astype(np.float32)).astype(np.float32)


def _weights(f):
    """saturate_cast<short>(cvRound(w * 2048)) for the two taps (1 - f, f); cvRound rounds half to even."""
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    return w0, w1


def linear_tables_x(src: int, dst: int):
    """Horizontal taps: outside the image the position is clamped AND the fraction reset (sx < 0 -> sx = 0, fx = 0;
    sx >= src - 1 -> sx = src - 1, fx = 0), so border columns copy the border pixel."""
    s, f = _positions(src, dst)
    lo, hi = s < 0, s >= src - 1
    f = f.copy()
    f[lo | hi] = 0
    s = np.where(lo, 0, np.where(hi, src - 1, s)).astype(np.int32)
    w0, w1 = _weights(f)
    return s, np.minimum(s + 1, src - 1).astype(np.int32), w0, w1


def linear_tables_y(src: int, dst: int):
    """Vertical taps: the fraction is kept and the two ROW INDICES are clipped instead, so border rows blend the
    border row with itself using both weights (two separate >> 16 roundings -- not the same as copying it)."""
    s, f = _positions(src, dst)
    w0, w1 = _weights(f)
    return np.clip(s, 0, src - 1).astype(np.int32), np.clip(s + 1, 0, src - 1).astype(np.int32), w0, w1


def resize_linear_u8(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv2.resize(img, (width, height)) for a uint8 HWC image (INTER_LINEAR): horizontal pass in int32 with 11-bit
    weights, vertical pass `(((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2`."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    sx0, sx1, a0, a1 = linear_tables_x(w, width)
    sy0, sy1, b0, b1 = linear_tables_y(h, height)
    s = img.astype(np.int32)
    rows = s[:, sx0, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]
    r0, r1 = rows[sy0], rows[sy1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def nearest_offsets(src: int, dst: int) -> np.ndarray:
    """resizeNN: sx = min(cvFloor(dx * (1 / (dst / src))), src - 1)."""
    ifx = 1.0 / (dst / src)
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1).astype(np.int32)


def resize_nearest(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv2.resize(img, (width, height), interpolation=cv2.INTER_NEAREST) for an HW or HWC array."""
    return img[nearest_offsets(img.shape[0], height)][:, nearest_offsets(img.shape[1], width)]
