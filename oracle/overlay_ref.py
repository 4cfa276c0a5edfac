"""STOM visual-prompt overlay oracle (numpy, integer arithmetic).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates, without PIL/cv2:
  * ``Image.alpha_composite(frame.convert("RGBA"), layer).convert("RGB")``
    as used at /root/reference/model/STOM.py:84-87, :157-160, :204-207 and
    /root/reference/utils/visual_prompt_generator.py:361-363 (Pillow's integer
    Porter-Duff "over" for an opaque destination);
  * ``STOM.warp``      /root/reference/model/STOM.py:145-160
  * ``STOM.warp_point`` circle stamping  /root/reference/model/STOM.py:163-207
  * ``draw_rectangle`` /root/reference/utils/visual_prompt_generator.py:102-104
Pinned against Pillow / cv2 / the reference's own STOM.warp run in this
container: tests/golden/make_golden.py -> tests/golden/overlay_*.npz.
"""
from __future__ import annotations

import numpy as np


def alpha_composite_ref(frame_rgb: np.ndarray, layer_rgba: np.ndarray) -> np.ndarray:
    """Pillow AlphaComposite.c with dst alpha == 255, then drop alpha.

    coef1 = a*128, coef2 = (255-a)*128, t = s*coef1 + d*coef2 + (0x80<<7),
    out = (((t>>8)+t)>>8)>>7 ; a == 0 -> out = d.
    """
    assert frame_rgb.dtype == np.uint8 and layer_rgba.dtype == np.uint8
    s = layer_rgba[..., :3].astype(np.uint32)
    a = layer_rgba[..., 3:4].astype(np.uint32)
    d = frame_rgb.astype(np.uint32)
    t = s * (a * 128) + d * ((255 - a) * 128) + (0x80 << 7)
    out = (((t >> 8) + t) >> 8) >> 7
    out = np.where(a == 0, d, out)
    return out.astype(np.uint8)


def warp_layer_ref(layer_rgba: np.ndarray, flow_x: float, flow_y: float) -> np.ndarray:
    """STOM.warp (/root/reference/model/STOM.py:145-155): forward-scatter every
    layer pixel with alpha > 0, in row-major order, to
    (int(y + flow_y), int(x + flow_x)); Python int() truncates toward zero;
    last writer wins; out-of-bounds dropped.  (The reference's argument names
    are swapped at the call site :133-135; here flow_x moves along x.)
    """
    h, w = layer_rgba.shape[:2]
    out = np.zeros_like(layer_rgba)
    ys, xs = np.nonzero(layer_rgba[:, :, 3] > 0)
    for y, x in zip(ys.tolist(), xs.tolist()):
        nx = int(x + flow_x)
        ny = int(y + flow_y)
        if 0 <= nx < w and 0 <= ny < h:
            out[ny, nx, :] = layer_rgba[y, x, :]
    return out


def shift_params_ref(flow: float, n: int):
    """Reduce a float flow along one axis of length n to the integer form the
    CUDA overlay uses: (shift, zero_extra).

    For x + flow >= 0, int(x+flow) = x + floor(flow) (one source per
    destination).  For -1 < x + flow < 0 truncation sends the source to 0 as
    well: destination 0 then has an extra, EARLIER source x0 = -1 - floor(flow)
    (zero_extra = 1).  Derived by evaluating the reference expression on every
    coordinate, so float64 rounding is honoured; raises if the map is not of
    this form.
    """
    idx = np.arange(n, dtype=np.int64)
    dst = np.array([int(v) for v in (idx + flow)], dtype=np.int64)  # STOM.py:151-152
    nonneg = (idx + flow) >= 0
    shifts = np.unique((dst - idx)[nonneg]) if nonneg.any() else np.array([int(np.floor(flow))])
    if shifts.size != 1:
        raise ValueError("flow does not reduce to an integer shift")
    shift = int(shifts[0])
    neg = (~nonneg) & (dst == 0)
    zero_extra = int(neg.any())
    if zero_extra:
        x0 = int(idx[neg][-1])
        if neg.sum() != 1 or x0 != -1 - shift:
            raise ValueError("unexpected truncation pattern")
    return shift, zero_extra


def warp_layer_gather_ref(layer_rgba, sx, zx, sy, zy):
    """Gather restatement of warp_layer_ref in the (shift, zero_extra) form --
    what the CUDA kernel evaluates per destination pixel.  Candidates for
    destination (ny, nx): rows {ny - sy} plus {-1 - sy} when ny == 0 and zy;
    same for columns.  The LAST candidate in row-major source order with
    alpha > 0 wins."""
    h, w = layer_rgba.shape[:2]
    out = np.zeros_like(layer_rgba)
    for ny in range(h):
        rows = [ny - sy]
        if ny == 0 and zy:
            rows = [-1 - sy, ny - sy]
        for nx in range(w):
            cols = [nx - sx]
            if nx == 0 and zx:
                cols = [-1 - sx, nx - sx]
            for y in rows:
                for x in cols:
                    if 0 <= y < h and 0 <= x < w and layer_rgba[y, x, 3] > 0:
                        out[ny, nx] = layer_rgba[y, x]
    return out


def box_mask_ref(h, w, box, width):
    """Pixel set of PIL ``ImageDraw.rectangle(box, outline=..., width=width)``
    (Pillow draw.c ImagingDrawRectangle, outline branch): for i < width the
    full rows t+i and b-i over [l..r], and the columns r-i and l+i between
    rows t+width (included) and b-width+1 (excluded), in either direction --
    which is what makes over-wide outlines spill; reproduced faithfully.  For
    2*width <= box size this is the inclusive rectangle minus the inclusive
    interior [l+width..r-width]x[t+width..b-width]."""
    l, t, r, b = [int(v) for v in box]
    width = max(int(width), 1)
    yy, xx = np.mgrid[0:h, 0:w]
    in_x = (xx >= min(l, r)) & (xx <= max(l, r))
    rows = in_x & (((yy >= t) & (yy < t + width)) | ((yy <= b) & (yy > b - width)))
    ya, yb = t + width, b - width + 1          # the line excludes its end point yb
    if ya < yb:
        in_y = (yy >= ya) & (yy < yb)
    elif ya > yb:
        in_y = (yy > yb) & (yy <= ya)
    else:
        in_y = np.zeros_like(yy, dtype=bool)   # zero-length line draws nothing
    cols = in_y & (((xx <= r) & (xx > r - width)) | ((xx >= l) & (xx < l + width)))
    return rows | cols


def box_layer_ref(h, w, box, width, rgba) -> np.ndarray:
    """draw_rectangle (visual_prompt_generator.py:102-104) on the transparent
    RGBA canvas of image_blending (:287-288)."""
    out = np.zeros((h, w, 4), dtype=np.uint8)
    out[box_mask_ref(h, w, box, width)] = np.asarray(rgba, dtype=np.uint8)
    return out


def circle_spans_ref(radius: int):
    """Half-widths of cv2.circle(..., thickness=FILLED) per row offset
    dy = -r..r (OpenCV's integer midpoint circle, drawing.cpp ``Circle``).
    Returns int array hw[2r+1]: row cy+dy covers cx-hw .. cx+hw inclusive."""
    r = int(radius)
    hw = np.full(2 * r + 1, -1, dtype=np.int64)
    err, dx, dy, plus, minus = 0, r, 0, 1, (r << 1) - 1
    while dx >= dy:
        # rows cy +- dy get half-width dx ; rows cy +- dx get half-width dy
        for (row, half) in ((dy, dx), (dx, dy)):
            for sgn in (-1, 1):
                i = r + sgn * row
                hw[i] = max(hw[i], half)
        dy += 1
        err += plus
        plus += 2
        mask = 0 if err <= 0 else -1  # (err <= 0) - 1
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return hw


def circle_layer_ref(h, w, cx, cy, radius, rgba) -> np.ndarray:
    """warp_point's stamp (/root/reference/model/STOM.py:195-201): filled
    cv2.circle of the prompt colour (alpha already clamped to [96,148], :174)."""
    out = np.zeros((h, w, 4), dtype=np.uint8)
    hw = circle_spans_ref(radius)
    for i, half in enumerate(hw.tolist()):
        y = cy - radius + i
        if half < 0 or not (0 <= y < h):
            continue
        x0, x1 = max(cx - half, 0), min(cx + half, w - 1)
        if x0 <= x1:
            out[y, x0:x1 + 1] = np.asarray(rgba, dtype=np.uint8)
    return out


def clamp_point_alpha_ref(rgba):
    """STOM.py:174: alpha = max(min(alpha, 148), 96)."""
    r, g, b, a = [int(v) for v in rgba]
    return (r, g, b, max(min(a, 148), 96))


def overlay_clip_ref(frames_u8, layer_rgba, per_frame):
    """Whole-clip overlay: frames [T,H,W,3] u8, one RGBA layer, and per-frame
    dicts {mode: 0 none | 1 shifted layer (sx,zx,sy,zy) | 2 circle (cx,cy,r,rgba)}
    -- the policy output of STOM.propagate_in_video (:72-141)."""
    out = frames_u8.copy()
    t, h, w, _ = frames_u8.shape
    for i, pf in enumerate(per_frame):
        mode = pf.get("mode", 0)
        if mode == 1:
            lay = warp_layer_gather_ref(layer_rgba, pf.get("sx", 0), pf.get("zx", 0), pf.get("sy", 0), pf.get("zy", 0))
            out[i] = alpha_composite_ref(frames_u8[i], lay)
        elif mode == 2:
            lay = circle_layer_ref(h, w, pf["cx"], pf["cy"], pf["r"], pf["rgba"])
            out[i] = alpha_composite_ref(frames_u8[i], lay)
    return out
