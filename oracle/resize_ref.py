"""Frame resize oracle: ``smart_resize`` + Pillow's bicubic ``Image.resize`` (numpy, integer arithmetic).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference resizes every frame before the processor through the third-party ``qwen_vl_utils``
(``process_vision_info``: /root/reference/app.py:296, :417, utils/dataset.py:76; pinned ``qwen_vl_utils==0.0.10`` in
/root/reference/requirements.txt:16, absent from this image).  Its published algorithm for a list of frames is
``fetch_image``: ``smart_resize(height, width, factor=28, min_pixels, max_pixels)`` then
``image.resize((resized_width, resized_height))`` -- Pillow's default BICUBIC filter on an RGB uint8 image.
``smart_resize`` is the same function HF ships (transformers ``image_processing_qwen2_vl.py:62-89``), importable here.
Pillow's resampler (src/libImaging/Resample.c: ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``) is restated below: per output coordinate a window of
``2*ceil(support*scale)+1`` double coefficients of the a=-0.5 cubic, normalised, converted to 22-bit fixed point with
round-half-away, horizontal pass then vertical pass with an intermediate uint8 image, each
``clip8((sum + 2^21) >> 22)``.
Pinned against ``PIL.Image.resize`` (Pillow 12.2.0 here; the reference pins 11.1.0, same algorithm) and HF's
``smart_resize`` in tests/test_resize_cpu.py and tests/golden/resize_pil.npz.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def smart_resize_ref(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56,
                     max_pixels: int = 14 * 14 * 4 * 1280):
    """transformers image_processing_qwen2_vl.py:62-89 (Python ``round`` = round-half-to-even)."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError("absolute aspect ratio must be smaller than 200")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs_ref(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box: (ksize, bounds [out,2], kk int32 [out,ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resampling pass of a uint8 [H,W,C] image along ``axis`` (0 = vertical, 1 = horizontal)."""
    in_size = img.shape[axis]
    ksize, bounds, kk = coeffs_ref(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_ref(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``Image.fromarray(img).resize((out_w, out_h))`` for an RGB uint8 [H,W,3] image: horizontal pass, then vertical
    (Resample.c ImagingResampleInner; a pass whose size does not change is skipped)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w = img.shape[:2]
    out = img
    if out_w != w:
        out = _pass(out, out_w, 1)
    if out_h != h:
        out = _pass(out, out_h, 0)
    return out.copy() if out is img else out
