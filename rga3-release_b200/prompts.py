"""Visual-prompt layers for the STOM overlay: the mask and scribble shapes of the reference's prompt generator,
rasterised on the GPU (csrc/raster.cu) with Pillow's exact pixel coverage, plus the generator's size / alpha rules.

Mirrors /root/reference/utils/visual_prompt_generator.py:
  draw_mask      :268-274  ImageDraw.polygon(coords, fill)              -> ``mask_layer`` / ``OverlaySpec.from_mask``
  draw_scribble  :230-252  1000*scale Bezier samples, ImageDraw.line(w) -> ``scribble_layer`` / ``OverlaySpec.from_scribble``
  image_blending :294-297  alpha ranges, :326-359 line widths scaled by max(W, H) / image_size_anchor
  get_bbox_from_mask :406-414
The random choices of the generator (colour, alpha, width, control points) are policy and stay with the caller: every
function takes them as arguments; the helpers below reproduce the ranges when a ``random.Random`` is supplied.
The layer is 1 byte per pixel (palette index 1 = the prompt colour), the cheapest form the overlay kernel reads.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .overlay import FrameOp, OverlaySpec


def prompt_alpha(shape: str, rng: Optional[random.Random] = None) -> int:
    """image_blending :294-297: alpha 188..224, 72..128 for the filled mask."""
    rng = rng or random
    return rng.randint(188, 224) if shape != "mask" else rng.randint(72, 128)


def prompt_line_width(shape: str, img_w: int, img_h: int, image_size_anchor: int = 336, width: Optional[int] = None,
                      visual_prompt_style: str = "", rng: Optional[random.Random] = None) -> int:
    """image_blending :326-359: the stroke width of each shape, scaled by max(W, H) / image_size_anchor; an explicit
    ``width`` is scaled the same way and floored at 1.  ('point' has a radius, not a width: see ``prompt_point_radius``.)"""
    rng = rng or random
    s = max(img_w, img_h) / image_size_anchor
    if shape == "rectangle":
        lw = max(int(3 * s), 1) if visual_prompt_style == "constant" else max(rng.randint(int(2 * s), int(8 * s)), 1)
    elif shape in ("ellipse", "triangle"):
        lw = max(rng.randint(int(2 * s), int(8 * s)), 1)
    elif shape == "arrow":
        lw = max(rng.randint(int(1 * s), int(6 * s)), 1)
    elif shape == "scribble":
        lw = max(rng.randint(int(12 * s), int(15 * s)), 1)
    elif shape == "mask":
        lw = rng.randint(int(0 * s), int(2 * s))
    elif shape == "mask contour":
        lw = max(rng.randint(int(1 * s), int(1.5 * s)), 1)
    else:
        raise ValueError(f"unknown prompt shape {shape!r}")
    return max(int(width * s), 1) if width is not None else lw


def get_bbox_from_mask(mask) -> Tuple[int, int, int, int]:
    """(left, top, right + 1, bottom + 1) of the non-zero pixels (:406-414); accepts numpy or torch (CPU / CUDA)."""
    m = torch.as_tensor(mask)
    rows = torch.nonzero(m.reshape(m.shape[0], -1).any(dim=1)).reshape(-1)
    cols = torch.nonzero(m.any(dim=0).reshape(-1)).reshape(-1)
    if rows.numel() == 0:
        raise IndexError("empty mask")          # the reference's np.where(...)[0][[0, -1]] raises on an empty mask as well
    return int(cols[0]), int(rows[0]), int(cols[-1]) + 1, int(rows[-1]) + 1


def scribble_points(ctrl: Sequence[Sequence[float]], img_w: int, img_h: int, image_size_anchor: int = 336) -> np.ndarray:
    """The int(1000 * max(W, H) / anchor) Bezier samples of draw_scribble (:244-246) for control points p0..p3."""
    n = int(1000 * max(img_w, img_h) / image_size_anchor)
    c = np.ascontiguousarray(np.asarray(ctrl, dtype=np.float64).reshape(8))
    out = np.empty((n, 2), dtype=np.float64)
    _lib.check(_lib.lib().b200vit_scribble_points(c.ctypes.data_as(C.POINTER(C.c_double)), n,
                                                  out.ctypes.data_as(C.POINTER(C.c_double))), "scribble_points")
    return out


def _layer(h: int, w: int, device, out: Optional[torch.Tensor]) -> torch.Tensor:
    if out is None:
        return torch.zeros((h, w), dtype=torch.uint8, device=device)
    if tuple(out.shape) != (h, w) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous CUDA uint8 tensor of shape {(h, w)}")
    return out


def mask_layer(segmentation: Sequence[Sequence[float]], h: int, w: int, device="cuda", index: int = 1,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """draw_mask: every contour ``[x0, y0, x1, y1, ...]`` of ``segmentation`` filled as ImageDraw.polygon does.
    Returns the uint8 [h, w] palette layer (``index`` where painted); pass ``out`` to draw onto an existing layer."""
    layer = _layer(h, w, device, out)
    counts = np.asarray([len(seg) // 2 for seg in segmentation], dtype=np.int32)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(seg, dtype=np.float64).reshape(-1)[: 2 * (len(seg) // 2)]
                                                for seg in segmentation]) if len(segmentation) else np.zeros(0))
    with torch.cuda.device(layer.device):
        rc = _lib.lib().b200vit_raster_polygons(flat.ctypes.data_as(C.POINTER(C.c_double)),
                                                counts.ctypes.data_as(C.POINTER(C.c_int32)), len(counts), h, w, index,
                                                layer.data_ptr(), torch.cuda.current_stream(layer.device).cuda_stream)
    _lib.check(rc, "raster_polygons")
    return layer


def lines_layer(points, width: int, h: int, w: int, device="cuda", index: int = 1,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One ImageDraw.line([p[i-1], p[i]], width=width) per consecutive pair of ``points`` [n, 2] (no joints)."""
    layer = _layer(h, w, device, out)
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 2))
    with torch.cuda.device(layer.device):
        rc = _lib.lib().b200vit_raster_lines(pts.ctypes.data_as(C.POINTER(C.c_double)), pts.shape[0], int(width), h, w, index,
                                             layer.data_ptr(), torch.cuda.current_stream(layer.device).cuda_stream)
    _lib.check(rc, "raster_lines")
    return layer


def scribble_layer(ctrl, width: int, h: int, w: int, device="cuda", image_size_anchor: int = 336, index: int = 1,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """draw_scribble: cubic Bezier through control points p0..p3, stroke ``width`` (already scaled, see
    ``prompt_line_width``)."""
    return lines_layer(scribble_points(ctrl, w, h, image_size_anchor), width, h, w, device, index, out)


def _spec(layer: torch.Tensor, rgb, alpha: int, ops: Sequence[FrameOp]) -> OverlaySpec:
    pal = np.zeros((2, 4), dtype=np.uint8)
    pal[1] = (int(rgb[0]), int(rgb[1]), int(rgb[2]), int(alpha))      # color_alpha = rgb_value + (alpha,)  (:299)
    return OverlaySpec.from_palette(layer, pal, ops, device=layer.device)


def overlay_from_mask(segmentation, rgb, alpha: int, h: int, w: int, ops: Sequence[FrameOp], device="cuda") -> OverlaySpec:
    """image_blending(shape='mask') as an OverlaySpec: transparent layer, contours filled with rgb + (alpha,)."""
    return _spec(mask_layer(segmentation, h, w, device), rgb, alpha, ops)


def overlay_from_scribble(ctrl, rgb, alpha: int, width: int, h: int, w: int, ops: Sequence[FrameOp], device="cuda",
                          image_size_anchor: int = 336) -> OverlaySpec:
    """image_blending(shape='scribble') as an OverlaySpec."""
    return _spec(scribble_layer(ctrl, width, h, w, device, image_size_anchor), rgb, alpha, ops)


OverlaySpec.from_mask = staticmethod(overlay_from_mask)
OverlaySpec.from_scribble = staticmethod(overlay_from_scribble)
