"""CPU oracle of the two image resizes in the reference's data mapper (map_10channel_case2, mask2former/utils/dataloader.py
:405-414) -- TEST INFRASTRUCTURE ONLY.  Both are third-party routines the reference calls and does not vendor:

* ``image_processor(images=[color, depth_colorful])`` resizes with Pillow: ``Image.resize((w, h), resample=BILINEAR)``
  (HF ``Mask2FormerImageProcessor`` PIL backend, ``resample = 2`` in checkpoints/standard/preprocessor_config.json).
  Restated from Pillow's ``ImagingResample`` (libImaging/Resample.c, 8 bits per channel): separable convolution, the
  triangle filter widened by the scale when shrinking (antialiasing), coefficients normalised in double precision and
  rounded to 22 fractional bits, each pass accumulates in int32 starting from 1 << 21 and clips to 0..255; horizontal
  pass first, then vertical, with a uint8 intermediate.
* ``cv2.resize(depth, (h, w), interpolation=cv2.INTER_LINEAR)`` on the single-channel uint8 depth: OpenCV's 8-bit linear
  path (imgproc/resize.cpp): float source coordinates ``(dx + 0.5) * scale - 0.5``, 11-bit fixed-point weights
  (``saturate_cast<short>(w * 2048)``), horizontal pass into int32 rows, vertical pass
  ``(((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2``.

Pinned by tests/golden/resize.npz = outputs of Pillow and OpenCV themselves (oracle/make_golden_resize.py).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bilinear (triangle, support 1) filter over the whole
    axis.  Returns (bounds (out,2) int32 = first source index and count, integer coefficients (out, ksize), ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            w = 1.0 - abs(a) if abs(a) < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    ik = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64))
    return bounds, ik.astype(np.int32), ksize


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_bilinear_resize_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """(H,W,C) or (H,W) uint8 -> (h,w[,C]) uint8 exactly as ``PIL.Image.resize((w,h), BILINEAR)``."""
    squeeze = img.ndim == 2
    src = img[:, :, None] if squeeze else img
    H, W, C = src.shape
    h, w = out_hw
    cur = src.astype(np.int64)
    if w != W:                                           # horizontal pass (Pillow skips a pass that keeps the size)
        bounds, kk, _ = pil_bilinear_coeffs(W, w)
        out = np.empty((H, w, C), dtype=np.uint8)
        for xx in range(w):
            x0, n = bounds[xx]
            acc = (cur[:, x0:x0 + n, :] * kk[xx, :n, None].astype(np.int64)).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            out[:, xx, :] = _clip8(acc)
        cur = out.astype(np.int64)
    if h != H:
        bounds, kk, _ = pil_bilinear_coeffs(H, h)
        out = np.empty((h, cur.shape[1], C), dtype=np.uint8)
        for yy in range(h):
            y0, n = bounds[yy]
            acc = (cur[y0:y0 + n, :, :] * kk[yy, :n, None, None].astype(np.int64)).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        cur = out.astype(np.int64)
    res = cur.astype(np.uint8)
    return res[:, :, 0] if squeeze else res


def cv_linear_tables(in_size: int, out_size: int, clamp_weights: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """OpenCV resize (INTER_LINEAR, 8U): source index and the two 11-bit weights per destination index.  Along x the
    weight is reset to (1, 0) where the window leaves the image; along y (``clamp_weights=False``) OpenCV keeps the
    fractional weights and clips the two ROW INDICES instead, so border rows blend a row with itself through two
    truncating products."""
    scale = 1.0 / (out_size / in_size)             # OpenCV: inv_scale = (double)dsize/ssize; scale = 1. / inv_scale
    ofs = np.zeros(out_size, dtype=np.int32)
    wts = np.zeros((out_size, 2), dtype=np.int32)
    for d in range(out_size):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_weights:
            if s < 0:
                f, s = np.float32(0.0), 0
            if s >= in_size - 1:
                f, s = np.float32(0.0), in_size - 1
        ofs[d] = s
        w0 = np.float32(np.float32(1.0) - f) * np.float32(2048.0)
        w1 = f * np.float32(2048.0)
        wts[d] = (int(np.rint(w0)), int(np.rint(w1)))      # saturate_cast<short>: round half to even
    return ofs, wts


def cv_linear_resize_u8(img: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """(H,W) uint8 -> (h,w) uint8 exactly as ``cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)``."""
    H, W = img.shape
    h, w = out_hw
    xofs, alpha = cv_linear_tables(W, w)
    yofs, beta = cv_linear_tables(H, h, clamp_weights=False)
    src = img.astype(np.int64)
    x1 = np.minimum(xofs + 1, W - 1)
    rows = src[:, xofs] * alpha[None, :, 0] + src[:, x1] * alpha[None, :, 1]          # (H, w) scaled by 2048
    r0, r1 = rows[np.clip(yofs, 0, H - 1)], rows[np.clip(yofs + 1, 0, H - 1)]
    b0, b1 = beta[:, 0][:, None].astype(np.int64), beta[:, 1][:, None].astype(np.int64)
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
