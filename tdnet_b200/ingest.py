"""Frame ingest and result export on either side of the model call, on the device.

The reference does both with OpenCV on the host (a dependency outside /root/reference, opencv-python==4.1.1.26):
  * Testing/dataloader.py:63   `cv2.resize(img, self.size)`: the uint8 RGB frame is resized (INTER_LINEAR, OpenCV's
    8-bit fixed-point path) before `/255`, mean / std and the NCHW transpose (:63-71);
  * Testing/test.py:61-64      `output.max(1)[1]` -> int8 -> `cv2.resize(pred, (W//4, H//4), INTER_NEAREST)`.
Here `FrameResizer` runs the first on the GPU bit-exactly (`tdn_resize_linear_u8`; its result feeds
`model.forward_u8`, which applies the normalisation inside the stem kernel), and `nearest_coords` gives the
full-resolution pixel coordinates the second one samples, so that `model.forward_preview` interpolates and arg-maxes
only those pixels (`tdn_upsample_argmax_sampled`).

OpenCV derives its interpolation taps with float / double host arithmetic; the functions below reproduce exactly that
(resize.cpp: `fx = (float)((dx + 0.5) * scale - 0.5)`, `cvFloor`, `cvRound(w * 2048)`), the kernels only do the
integer part.  tests/test_ingest.py checks the tables against the oracle restatement and against cv2 itself.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

_COEF_SCALE = 2048.0   # 1 << INTER_RESIZE_COEF_BITS


def linear_taps(src: int, dst: int, axis: str) -> np.ndarray:
    """int32 [dst, 4] = {offset0, offset1, weight0, weight1} of OpenCV's 8-bit INTER_LINEAR for one axis.
    axis 'x': positions outside the image are clamped and their fraction reset (the border pixel is copied);
    axis 'y': the fraction is kept and the two row indices are clipped (the border row is blended with itself)."""
    if src < 1 or dst < 1:
        raise ValueError("linear_taps: empty axis")
    scale = 1.0 / (dst / src)
    f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if axis == "x":
        out = (s < 0) | (s >= src - 1)
        f[out] = 0
        s0 = np.clip(s, 0, src - 1)
        s1 = np.minimum(s0 + 1, src - 1)
    elif axis == "y":
        s0, s1 = np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1)
    else:
        raise ValueError("axis must be 'x' or 'y'")
    w0 = np.rint((np.float32(1.0) - f) * np.float32(_COEF_SCALE))
    w1 = np.rint(f * np.float32(_COEF_SCALE))
    return np.stack([s0, s1, w0.astype(np.int64), w1.astype(np.int64)], axis=1).astype(np.int32)


def nearest_coords(src: int, dst: int) -> np.ndarray:
    """int32 [dst]: the source index cv2.INTER_NEAREST reads for each output index, min(floor(d * src / dst), src - 1)
    with the scale computed as 1 / (dst / src) in double (resize.cpp, resizeNN)."""
    ifx = 1.0 / (dst / src)
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1).astype(np.int32)


class FrameResizer:
    """`cv2.resize(frame, (W, H))` (Testing/dataloader.py:63) for uint8 HWC RGB frames on the GPU, bit-exact.

        resize = FrameResizer((1024, 2048), (769, 1537), "cuda:0")
        logits = model.forward_u8(resize(frame_u8), pos_id)        # frame_u8: [n, 1024, 2048, 3] uint8 on the device
    """

    def __init__(self, src_hw, dst_hw, device):
        self.src_hw, self.dst_hw = tuple(src_hw), tuple(dst_hw)
        self.device = torch.device(device)
        self.lib = _cabi.load()
        self.x_taps = torch.from_numpy(linear_taps(self.src_hw[1], self.dst_hw[1], "x")).contiguous().to(self.device)
        self.y_taps = torch.from_numpy(linear_taps(self.src_hw[0], self.dst_hw[0], "y")).contiguous().to(self.device)

    def __call__(self, frames_u8: torch.Tensor) -> torch.Tensor:
        if not frames_u8.is_cuda:
            raise RuntimeError("tdnet_b200 runs on a CUDA (sm_100) device only; there is no CPU path.")
        if (frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3
                or tuple(frames_u8.shape[1:3]) != self.src_hw):
            raise RuntimeError(f"FrameResizer expects uint8 HWC frames [n, {self.src_hw[0]}, {self.src_hw[1]}, 3]")
        frames_u8 = frames_u8.contiguous()
        n = frames_u8.shape[0]
        out = torch.empty((n, self.dst_hw[0], self.dst_hw[1], 3), dtype=torch.uint8, device=frames_u8.device)
        stream = torch.cuda.current_stream(frames_u8.device).cuda_stream
        _cabi.check(self.lib.tdn_resize_linear_u8(frames_u8.data_ptr(), n, self.src_hw[0], self.src_hw[1],
                                                  self.x_taps.data_ptr(), self.y_taps.data_ptr(), out.data_ptr(),
                                                  self.dst_hw[0], self.dst_hw[1], C.c_void_p(stream)), "resize_linear_u8")
        return out
