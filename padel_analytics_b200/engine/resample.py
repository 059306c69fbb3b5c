"""Host-side coefficient tables for the two bit-exact resamplers the reference pipeline uses.

* cv2.resize(..., INTER_LINEAR) on uint8 — used by ultralytics LetterBox ([3P], reached from
  /root/reference/trackers/players_tracker/players_tracker.py:351).  OpenCV's 8-bit path works in 11-bit fixed
  point (INTER_RESIZE_COEF_BITS=11): per-axis integer coefficient pairs.
* PIL.Image.resize default (BICUBIC with antialiasing support scaling) — used by
  players_keypoints_tracker.py:260-266, keypoints_tracker.py:190-194 and ball_tracker/iterable.py:188.
  Pillow's 8bpc path normalises float64 coefficients to 22-bit fixed point (PRECISION_BITS = 32-8-2).

Only the tables are computed here (float math identical to the libraries'); the per-pixel integer arithmetic
runs in csrc/preprocess.cu.
"""
from __future__ import annotations

import math

import numpy as np


# ---------------------------------------------------------------------------------------------------------
# OpenCV INTER_LINEAR (8-bit, fixed point)
# ---------------------------------------------------------------------------------------------------------
def cv2_linear_tables(src: int, dst: int):
    """Return (ofs int32[dst], coef int32[dst,2]) for one axis, as cv::resize computes them:
    fx = float((d+0.5)*scale - 0.5), s = floor(fx), fx -= s, clamped at the borders; coefficients are
    saturate_cast<short>(c * 2048) with round-half-even."""
    scale = 1.0 / (float(dst) / float(src))  # double, like cv::resize (scale_x = 1./inv_scale_x)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0.0
    hi = s >= src - 1
    s[hi] = src - 1
    f[hi] = 0.0
    c0 = np.clip(np.rint((np.float32(1.0) - f) * np.float32(2048.0)), -32768, 32767).astype(np.int32)
    c1 = np.clip(np.rint(f * np.float32(2048.0)), -32768, 32767).astype(np.int32)
    return s.astype(np.int32), np.stack([c0, c1], axis=1).astype(np.int32)


def letterbox_geometry(h: int, w: int, imgsz: int, stride: int = 32, auto: bool = True):
    """ultralytics LetterBox(new_shape=(imgsz,imgsz), auto=, stride=, scaleup=True, center=True) geometry
    ([3P]; SURVEY App. A.4 ii).  Returns dict(Hn, Wn, rh, rw, top, left)."""
    nh = nw = imgsz
    r = min(nh / h, nw / w)
    rw, rh = int(round(w * r)), int(round(h * r))
    dw, dh = nw - rw, nh - rh
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return dict(Hn=rh + top + bottom, Wn=rw + left + right, rh=rh, rw=rw, top=top, left=left)


# ---------------------------------------------------------------------------------------------------------
# Pillow BICUBIC with antialias (ImagingResample, 8bpc)
# ---------------------------------------------------------------------------------------------------------
def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x: float) -> float:
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


def pil_bilinear_tables(in_size: int, out_size: int):
    """Same for Image.BILINEAR (support 1): torchvision transforms.Resize on a PIL image
    (/root/reference/trackers/keypoints_tracker/iterable.py:19)."""
    return pil_bicubic_tables(in_size, out_size, support0=1.0, filt=_bilinear)


def pil_bicubic_tables(in_size: int, out_size: int, support0: float = 2.0, filt=None):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
    Returns (bounds int32[out,2] = (xmin, xsize), kk int32[out,ksize], ksize)."""
    filt = filt or _bicubic
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = filt((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds[xx, 0] = xmin
        bounds[xx, 1] = xmax
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
    return bounds, kk, ksize
