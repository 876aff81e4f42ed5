"""CPU oracle for the crop resize of the hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates what ``cv2.resize(face, (224, 224), interpolation=cv2.INTER_AREA)`` followed by
``cv2.cvtColor(face, cv2.COLOR_RGB2BGR)`` computes at the reference call sites
(/root/reference/CViT-main/cvit_prediction.py:114-115, also :96-97, :141-142 and
preprocessing/extractfaces.py:129-130) for uint8 HWC crops.

The arithmetic lives in OpenCV (third-party, not under /root/reference, no version pinned by
the reference — no requirements file exists; the image has opencv-python 4.13.0).  This file
restates OpenCV's published algorithm (modules/imgproc/src/resize.cpp):

* both scale factors >= 1 and integer  -> ``resizeAreaFast_``: block sum * (1/area) in fp32,
  round-half-even; the 2x2 case uses OpenCV's SIMD formula (a+b+c+d+2)>>2;
* both scale factors >= 1, fractional  -> ``resizeArea_`` with the ``computeResizeAreaTab``
  weights (fp64 -> fp32), fp32 accumulation in table order, round-half-even;
* otherwise (any up-scaling)           -> the bilinear kernel in "area mode": 11-bit fixed-point
  coefficients, ``HResizeLinear`` / ``VResizeLinear`` integer arithmetic.

Pinned in tests/test_resize_oracle.py against the installed cv2 on a sweep of crop sizes
(bit-exact on the linear and integer paths; <= 1 LSB on the fractional-area path, where
OpenCV's SIMD build may contract multiply-adds differently).
"""
from __future__ import annotations

import math

import numpy as np

OUT = 224
COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _area_tab(ssize: int, dsize: int, scale: float):
    """computeResizeAreaTab: list of (di, si, alpha fp32) in OpenCV's order."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _round_half_even_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def _resize_area_frac(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw, cn = src.shape
    xtab = _area_tab(sw, dw, sw / dw)
    ytab = _area_tab(sh, dh, sh / dh)
    s = src.astype(np.float32)
    # horizontal pass per source row, in table order (fp32, separate multiply and add)
    di = np.array([t[0] for t in xtab]); si = np.array([t[1] for t in xtab])
    al = np.array([t[2] for t in xtab], dtype=np.float32)
    out = np.zeros((dh, dw, cn), dtype=np.float32)
    started = np.zeros(dh, dtype=bool)
    for (dy, sy, beta) in ytab:
        buf = np.zeros((dw, cn), dtype=np.float32)
        row = s[sy]
        # sequential accumulation in table order: entries of one dx are consecutive
        for k in range(len(xtab)):
            buf[di[k]] = buf[di[k]] + row[si[k]] * al[k]
        if not started[dy]:
            out[dy] = beta * buf
            started[dy] = True
        else:
            out[dy] = out[dy] + beta * buf
    return _round_half_even_u8(out)


def _resize_area_fast(src: np.ndarray, dw: int, dh: int, isx: int, isy: int) -> np.ndarray:
    sh, sw, cn = src.shape
    blk = src[:dh * isy, :dw * isx].astype(np.int64).reshape(dh, isy, dw, isx, cn).sum(axis=(1, 3))
    if isx == 2 and isy == 2:
        return ((blk + 2) >> 2).astype(np.uint8)
    scale = np.float32(1.0 / (isx * isy))
    return _round_half_even_u8(blk.astype(np.float32) * scale)


def _linear_coeffs(ssize: int, dsize: int):
    """Area-mode bilinear taps: (sx, a0, a1 int16-like, is_right_edge)."""
    scale = ssize / dsize
    inv = dsize / ssize
    sxs, a0s, a1s = [], [], []
    xmax = dsize
    for dx in range(dsize):
        sx = math.floor(dx * scale)
        fx = np.float32((dx + 1) - (sx + 1) * inv)
        fx = np.float32(0.0) if fx <= 0 else np.float32(fx - math.floor(fx))
        if sx < 0:
            fx, sx = np.float32(0.0), 0
        if sx + 1 >= ssize:
            xmax = min(xmax, dx)
            if sx >= ssize - 1:
                fx, sx = np.float32(0.0), ssize - 1
        c0 = np.float32(1.0) - fx
        a0 = int(np.clip(np.rint(np.float32(c0 * np.float32(COEF_SCALE))), -32768, 32767))
        a1 = int(np.clip(np.rint(np.float32(fx * np.float32(COEF_SCALE))), -32768, 32767))
        sxs.append(sx); a0s.append(a0); a1s.append(a1)
    return np.array(sxs), np.array(a0s, dtype=np.int64), np.array(a1s, dtype=np.int64), xmax


def _resize_linear_area_mode(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw, cn = src.shape
    sx, a0, a1, xmax = _linear_coeffs(sw, dw)
    sy, b0, b1, _ = _linear_coeffs(sh, dh)
    s = src.astype(np.int64)
    sx1 = np.minimum(sx + 1, sw - 1)
    rows = s[:, sx, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]     # HResizeLinear
    edge = np.arange(dw) >= xmax
    rows[:, edge, :] = s[:, sx[edge], :] * COEF_SCALE
    sy1 = np.minimum(sy + 1, sh - 1)
    r0 = rows[sy]; r1 = rows[sy1]
    v = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2   # VResizeLinear
    return np.clip(v, 0, 255).astype(np.uint8)


def resize_area_u8(src: np.ndarray, dw: int = OUT, dh: int = OUT) -> np.ndarray:
    """uint8 [h,w,3] -> uint8 [dh,dw,3], cv2.INTER_AREA semantics."""
    sh, sw, _ = src.shape
    scale_x, scale_y = sw / dw, sh / dh
    if scale_x >= 1 and scale_y >= 1:
        isx, isy = int(round(scale_x)), int(round(scale_y))
        if abs(scale_x - isx) < np.finfo(np.float64).eps and abs(scale_y - isy) < np.finfo(np.float64).eps:
            return _resize_area_fast(src, dw, dh, isx, isy)
        return _resize_area_frac(src, dw, dh)
    return _resize_linear_area_mode(src, dw, dh)


def crop_to_model_input(face: np.ndarray, swap_rb: bool = True) -> np.ndarray:
    """cvit_prediction.py:114-115: resize to 224x224 INTER_AREA, then RGB<->BGR swap."""
    out = resize_area_u8(face, OUT, OUT)
    return out[:, :, ::-1].copy() if swap_rb else out
