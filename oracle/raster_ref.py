"""CPU restatement of Pillow's polygon fill and wide line (test infrastructure, never imported by the product).

The mask and scribble prompts of /root/reference/utils/visual_prompt_generator.py are drawn with
``ImageDraw.polygon(coords, fill=...)`` (draw_mask :268-274) and ``ImageDraw.line([prev, cur], width=...)``
(draw_scribble :230-252).  The pixel coverage is decided inside Pillow's C library (libImaging/Draw.c; the reference
pins Pillow 11.1.0, this image has 12.2.0), which is not part of /root/reference: the functions below restate
ImagingDrawPolygon (fill branch), polygon_generic, ImagingDrawWideLine, line32 and hline32, with float32 / float64
arithmetic placed where the C code has it.  Pinned by fuzzing against the live library (tests/test_raster_cpu.py) and by
the golden layers the reference's own draw_mask / draw_scribble produced (tests/golden/prompt_layers.npz).
Pure-Python loops: for small cases only.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def _roundf(v) -> float:
    """C roundf: halves away from zero."""
    v = float(v)
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def round_up_f(f) -> int:
    """ROUND_UP on a float32 argument: `(f) + 0.5F` stays float32 for f >= 0, the other branch is double."""
    f = f32(f)
    return int(math.floor(f32(f + f32(0.5)))) if f >= 0 else -int(math.floor(float(abs(f)) + 0.5))


def round_down_f(f) -> int:
    f = f32(f)
    return int(math.ceil(f32(f - f32(0.5)))) if f >= 0 else -int(math.ceil(float(abs(f)) - 0.5))


def round_up_d(f: float) -> int:
    return int(math.floor(f + 0.5)) if f >= 0.0 else -int(math.floor(abs(f) + 0.5))


def round_down_d(f: float) -> int:
    return int(math.ceil(f - 0.5)) if f >= 0.0 else -int(math.ceil(abs(f) - 0.5))


class Edge:
    __slots__ = ("x0", "y0", "xmin", "ymin", "xmax", "ymax", "dx")

    def __init__(self, x0, y0, x1, y1):      # add_edge
        self.xmin, self.xmax = min(x0, x1), max(x0, x1)
        self.ymin, self.ymax = min(y0, y1), max(y0, y1)
        self.dx = f32(0.0) if y0 == y1 else f32(f32(x1 - x0) / f32(y1 - y0))
        self.x0, self.y0 = x0, y0

    def x_at(self, y):                        # (ymin - y0) * dx + x0 in float32, two roundings
        return f32(f32(f32(y - self.y0) * self.dx) + f32(self.x0))


def hline(im, x0, y, x1):
    h, w = im.shape
    if 0 <= y < h:
        x0, x1 = max(x0, 0), min(x1, w - 1)
        if x0 <= x1:
            im[y, x0:x1 + 1] = 1


def polygon_generic(im, edges):
    """Scanline fill of Pillow's polygon_generic for a non-blending ink (hline32)."""
    h, w = im.shape
    if not edges:
        return
    ymin, ymax = h - 1, 0
    table = []
    for e in edges:
        ymin, ymax = min(ymin, e.ymin), max(ymax, e.ymax)
        if e.ymin == e.ymax:
            hline(im, e.xmin, e.ymin, e.xmax)
            continue
        table.append(e)
    ymin, ymax = max(ymin, 0), min(ymax, h)
    for y in range(ymin, ymax + 1):
        xx = []
        for i, cur in enumerate(table):
            if not (cur.ymin <= y <= cur.ymax):
                continue
            x = cur.x_at(y)
            if y == cur.ymax and y < ymax:
                xx += [x, x]                                  # "needed to draw consistent polygons"
                continue
            if (y == cur.ymin or y == cur.ymax) and cur.dx != 0:
                adj = y - 1 if y == cur.ymax else y + 1       # "connect discontiguous corners"
                for k in range(i):
                    o = table[k]
                    if (y != o.ymin and y != o.ymax) or o.dx == 0:
                        continue
                    if _roundf(x) != _roundf(o.x_at(y)):
                        continue
                    if adj < o.ymin or adj > o.ymax:
                        continue
                    ax, ao = cur.x_at(adj), o.x_at(adj)
                    if x > f32(ax + f32(1)) and x > f32(ao + f32(1)):
                        x = f32(f32(_roundf(max(ax, ao))) + f32(1))
                    elif f32(ax - f32(1)) > x and f32(ao - f32(1)) > x:
                        x = f32(f32(_roundf(min(ax, ao))) - f32(1))
                    break
            xx.append(x)
        xx.sort()
        for i in range(1, len(xx), 2):
            hline(im, round_up_f(xx[i - 1]), y, round_down_f(xx[i]))


def polygon_edges(xy):
    """ImagingDrawPolygon, fill branch: xy = [(x, y), ...] as ImageDraw receives them (floats are truncated)."""
    pts = [(int(x), int(y)) for x, y in xy]
    edges = []
    for i in range(len(pts) - 1):
        (x0, y0), (x1, y1) = pts[i], pts[i + 1]
        if y0 == y1 and i != 0 and y0 == pts[i - 1][1] and edges:
            if x1 > x0 > pts[i - 1][0]:
                edges[-1].xmax = x1
                continue
            if x1 < x0 < pts[i - 1][0]:
                edges[-1].xmin = x1
                continue
        edges.append(Edge(x0, y0, x1, y1))
    if pts and pts[-1] != pts[0]:
        edges.append(Edge(pts[-1][0], pts[-1][1], pts[0][0], pts[0][1]))
    return edges


def fill_polygon_ref(im, xy):
    """coverage of ImageDraw.polygon(xy, fill=ink, outline=None) on a uint8 [h, w] array (1 = painted)."""
    polygon_generic(im, polygon_edges(xy))


def _line32(im, x0, y0, x1, y1):
    h, w = im.shape

    def pt(x, y):
        if 0 <= x < w and 0 <= y < h:
            im[y, x] = 1
    dx, dy = x1 - x0, y1 - y0
    xs, ys = (1 if dx >= 0 else -1), (1 if dy >= 0 else -1)
    dx, dy = abs(dx), abs(dy)
    if dx == 0:
        for _ in range(dy):
            pt(x0, y0)
            y0 += ys
    elif dy == 0:
        for _ in range(dx):
            pt(x0, y0)
            x0 += xs
    elif dx > dy:
        n = dx
        dy += dy
        e = dy - dx
        dx += dx
        for _ in range(n):
            pt(x0, y0)
            if e >= 0:
                y0 += ys
                e -= dx
            e += dy
            x0 += xs
    else:
        n = dy
        dx += dx
        e = dx - dy
        dy += dy
        for _ in range(n):
            pt(x0, y0)
            if e >= 0:
                x0 += xs
                e -= dy
            e += dx
            y0 += ys


def line_ref(im, p0, p1, width):
    """coverage of ImageDraw.line([p0, p1], fill=ink, width=width) (two points, no joint)."""
    h, w = im.shape
    x0, y0, x1, y1 = int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1])
    if width <= 1:
        _line32(im, x0, y0, x1, y1)
        if 0 <= x1 < w and 0 <= y1 < h:
            im[y1, x1] = 1                                     # "draw last point"
        return
    dx, dy = x1 - x0, y1 - y0
    if dx == 0 and dy == 0:
        if 0 <= x0 < w and 0 <= y0 < h:
            im[y0, x0] = 1
        return
    big = math.hypot(dx, dy)                                   # ImagingDrawWideLine
    small = (width - 1) / 2.0
    ratio_max, ratio_min = round_up_d(small) / big, round_down_d(small) / big
    dxmin, dxmax = round_down_d(ratio_min * dy), round_down_d(ratio_max * dy)
    dymin, dymax = round_up_d(ratio_min * dx), round_up_d(ratio_max * dx)
    v = [(x0 - dxmin, y0 + dymax), (x1 - dxmin, y1 + dymax), (x1 + dxmax, y1 - dymin), (x0 + dxmax, y0 - dymin)]
    polygon_generic(im, [Edge(*v[i], *v[(i + 1) % 4]) for i in range(4)])


def scribble_points_ref(ctrl, n_points):
    """The Bezier samples of draw_scribble (:244-246), evaluated with the reference's own Python expression."""
    p0, p1, p2, p3 = ctrl
    out = []
    for t in np.linspace(0, 1, n_points):
        x = (1 - t)**3 * p0[0] + 3 * (1 - t)**2 * t * p1[0] + 3 * (1 - t) * t**2 * p2[0] + t**3 * p3[0]
        y = (1 - t)**3 * p0[1] + 3 * (1 - t)**2 * t * p1[1] + 3 * (1 - t) * t**2 * p2[1] + t**3 * p3[1]
        out.append((float(x), float(y)))
    return np.asarray(out, dtype=np.float64)


def mask_layer_ref(segmentation, h, w):
    """draw_mask (:268-274): every contour filled; 1 where painted."""
    im = np.zeros((h, w), dtype=np.uint8)
    for seg in segmentation:
        fill_polygon_ref(im, [(seg[i], seg[i + 1]) for i in range(0, len(seg), 2)])
    return im


def scribble_layer_ref(ctrl, width, h, w, image_size_anchor=336):
    """draw_scribble (:230-252): int(1000 * max(w, h) / anchor) samples, one wide line per step."""
    pts = scribble_points_ref(ctrl, int(1000 * max(w, h) / image_size_anchor))
    im = np.zeros((h, w), dtype=np.uint8)
    for i in range(1, len(pts)):
        line_ref(im, pts[i - 1], pts[i], width)
    return im
