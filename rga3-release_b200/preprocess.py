"""Frame sizing ahead of the visual path: ``smart_resize`` + bicubic resize of uint8 frames on the GPU.

Mirrors what the reference does to every frame through ``qwen_vl_utils.process_vision_info``
(/root/reference/app.py:296, :417, utils/dataset.py:76): ``fetch_image`` computes
``smart_resize(height, width, factor=28, min_pixels, max_pixels)`` and calls ``image.resize((w, h))`` (Pillow BICUBIC).
The pixel work runs in csrc/resize.cu, bit-identical to Pillow; the arithmetic of ``smart_resize`` is host integer math
(transformers image_processing_qwen2_vl.py:62-89).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib

IMAGE_FACTOR = 28
MIN_PIXELS = 4 * 28 * 28          # qwen_vl_utils defaults
MAX_PIXELS = 16384 * 28 * 28


def smart_resize(height: int, width: int, factor: int = IMAGE_FACTOR, min_pixels: int = MIN_PIXELS,
                 max_pixels: int = MAX_PIXELS, variant: str = "qwen_vl_utils") -> Tuple[int, int]:
    """(h_bar, w_bar): both multiples of ``factor``, pixel count within [min_pixels, max_pixels], aspect ratio kept as
    closely as possible (Python ``round`` is round-half-to-even, as in both sources).

    ``variant="qwen_vl_utils"`` (default) is what the reference runs on every frame: qwen_vl_utils 0.0.10
    (requirements.txt:16) ``vision_process.smart_resize`` -- the first rounding is floored at ``factor``
    (``max(factor, round_by_factor(...))``: a side shorter than 14 px does not collapse to 0) and the max_pixels branch
    is a bare ``floor_by_factor``.  The package is not installed here; restated from its published source.
    ``variant="transformers"`` is the HF processor's formula (image_processing_qwen2_vl.py:62-89): no floor on the first
    rounding, ``max(factor, ...)`` in the max_pixels branch.  The two differ only for very small or extreme-aspect
    frames, e.g. 10 x 1000 -> (28, 1008) vs (28, 560)."""
    if height <= 0 or width <= 0:
        raise ValueError("height and width must be positive")
    if max(height, width) / min(height, width) > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
    if variant == "qwen_vl_utils":
        h_bar = max(factor, round(height / factor) * factor)
        w_bar = max(factor, round(width / factor) * factor)
        if h_bar * w_bar > max_pixels:
            beta = math.sqrt((height * width) / max_pixels)
            h_bar = math.floor(height / beta / factor) * factor
            w_bar = math.floor(width / beta / factor) * factor
        elif h_bar * w_bar < min_pixels:
            beta = math.sqrt(min_pixels / (height * width))
            h_bar = math.ceil(height * beta / factor) * factor
            w_bar = math.ceil(width * beta / factor) * factor
        return h_bar, w_bar
    if variant != "transformers":
        raise ValueError("variant must be 'qwen_vl_utils' or 'transformers'")
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


def resize_frames(frames_u8: torch.Tensor, out_h: int, out_w: int, out: Optional[torch.Tensor] = None,
                  stream=None) -> torch.Tensor:
    """``[T,H,W,3]`` uint8 CUDA frames -> ``[T,out_h,out_w,3]``, each frame exactly
    ``PIL.Image.fromarray(f).resize((out_w, out_h))``."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
        raise ValueError("frames must be a CUDA uint8 tensor [T,H,W,3]")
    fr = frames_u8.contiguous()
    t, h, w, _ = fr.shape
    if out is None:
        out = torch.empty((t, out_h, out_w, 3), dtype=torch.uint8, device=fr.device)
    elif tuple(out.shape) != (t, out_h, out_w, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError(f"out must be a contiguous CUDA uint8 tensor of shape {(t, out_h, out_w, 3)}")
    l = _lib.lib()
    ws_bytes = int(l.b200vit_resize_workspace_bytes(t, h, w, out_h, out_w))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=fr.device)
    s = stream if stream is not None else torch.cuda.current_stream(fr.device).cuda_stream
    with torch.cuda.device(fr.device):
        rc = l.b200vit_resize_bicubic(fr.data_ptr(), t, h, w, out.data_ptr(), out_h, out_w, ws.data_ptr(), ws_bytes, s)
    _lib.check(rc, "resize")
    out._keep = (fr, ws)  # inputs stay alive until the enqueued kernels have run
    return out


def fit_frames(frames_u8: torch.Tensor, min_pixels: int = MIN_PIXELS, max_pixels: int = MAX_PIXELS) -> torch.Tensor:
    """smart_resize + resize: what ``fetch_image`` does to each frame of a clip."""
    t, h, w, _ = frames_u8.shape
    oh, ow = smart_resize(h, w, IMAGE_FACTOR, min_pixels, max_pixels)
    return resize_frames(frames_u8, oh, ow)
