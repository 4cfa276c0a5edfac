"""STOM placement-policy oracle (numpy; no cv2, no PIL).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates what /root/reference/model/STOM.py does between the tracker output and
the per-frame overlay:
  * ``propagate_in_video`` :72-141  -- which frame gets which overlay;
  * the MAD filter on the flows of the visible tracks :104-131 (numpy float32
    arithmetic: ``np.linalg.norm``, ``np.median``, ``np.mean`` with numpy's
    pairwise summation);
  * ``warp_point`` :163-207 -- visible-track mask, ``cv2.morphologyEx(MORPH_CLOSE)``
    with ``cv2.getStructuringElement(MORPH_ELLIPSE, (k, k))``, ``cv2.moments``
    centroid, circle radius and colour.
OpenCV is a third-party dependency of the reference (opencv-python, unpinned in
/root/reference/requirements.txt; 4.x installed here); its published algorithms
are restated below:
  * getStructuringElement(MORPH_ELLIPSE): r = k/2, c = k/2, row i (dy = i - r) holds
    ones in [max(c-dx,0), min(c+dx+1,k)) with dx = cvRound(c*sqrt((r*r-dy*dy)/(r*r)));
  * dilate: dst(y,x) = max, erode: dst(y,x) = min, both over the element's ones (i,j) of
    src(y+i-a, x+j-a) with anchor a = k/2 (the same correlation form for both -- checked
    against cv2 for even k, where it matters); out-of-image pixels are ignored;
  * moments of a uint8 image: exact integer sums, returned as doubles.
Pinned against the reference's own ``STOM.propagate_in_video`` (tracker stubbed) and
cv2 run in this container: tests/golden/make_golden.py -> tests/golden/stom_policy.npz.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from . import overlay_ref as ov


def pairwise_sum_f32(a: np.ndarray) -> np.float32:
    """numpy's float32 ``add.reduce`` over a 1-D array (numpy/_core/src/umath/loops_utils.h
    ``pairwise_sum``): n < 8 sequential; n <= 128 eight running sums over blocks of 8 combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail sequentially; larger n split at n/2 rounded down to
    a multiple of 8.  Every operation rounds to float32."""
    a = np.asarray(a, dtype=np.float32)
    n = a.shape[0]
    f = np.float32
    if n < 8:
        # numpy starts from -0.0 so that the sum of an empty / all -0.0 array keeps its sign
        res = f(-0.0)
        for i in range(n):
            res = f(res + a[i])
        return res
    if n <= 128:
        r = [f(a[i]) for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] = f(r[k] + a[i + k])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def mean_f32(a: np.ndarray) -> np.float32:
    """``np.mean`` of a float32 vector: pairwise float32 sum, then one float32 division."""
    return np.float32(pairwise_sum_f32(a) / np.float32(a.shape[0]))


def median_f32(a: np.ndarray) -> np.float32:
    """``np.median`` of a float32 vector: NaN if any NaN, else the middle order statistic (odd n) or the
    float32 mean of the two middle ones (their float32 sum divided by 2)."""
    a = np.sort(np.asarray(a, dtype=np.float32))
    n = a.shape[0]
    if np.isnan(a[-1]):
        return np.float32(np.nan)
    if n % 2 == 1:
        return np.float32(a[n // 2])
    return np.float32(np.float32(a[n // 2 - 1] + a[n // 2]) / np.float32(2))


def flow_shift_ref(key_track: np.ndarray, track: np.ndarray, vis: np.ndarray):
    """STOM.py:104-131.  Returns None (frame left untouched) or (flow_x, flow_y) as float32 -- the means the
    reference passes to ``warp`` (its ``avg_flow_y`` is the x flow, :126-135)."""
    vis = np.asarray(vis).astype(bool)
    f = np.float32
    flows = (track[vis].astype(f) - key_track[vis].astype(f)).astype(f)
    if len(flows) == 0:
        return None
    mag = np.sqrt((flows[:, 0] * flows[:, 0] + flows[:, 1] * flows[:, 1]).astype(f)).astype(f)
    med = median_f32(mag)
    mad = median_f32(np.abs((mag - med).astype(f)))
    thr = f(f(3) * mad)
    keep = (mag >= f(med - thr)) & (mag <= f(med + thr))
    filt = flows[keep]
    if len(filt) < vis.shape[0] // 2:
        return None
    if filt.size == 0:
        return f(0.0), f(0.0)
    fx, fy = mean_f32(filt[:, 0]), mean_f32(filt[:, 1])
    if np.isnan(fx) or np.isnan(fy):
        return None
    return fx, fy


def ellipse_rows_ref(k: int) -> List[Tuple[int, int]]:
    """cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)) as one [j1, j2) span per row (empty: j1 >= j2)."""
    r = k // 2
    c = k // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    rows = []
    for i in range(k):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * math.sqrt((r * r - dy * dy) * inv_r2)))  # cvRound: half to even
            rows.append((max(c - dx, 0), min(c + dx + 1, k)))
        else:
            rows.append((0, 0))
    return rows


def close_ref(mask: np.ndarray, k: int) -> np.ndarray:
    """cv2.morphologyEx(mask, MORPH_CLOSE, ellipse(k)) for a 0/255 mask (default anchor and border)."""
    h, w = mask.shape
    rows = ellipse_rows_ref(k)
    a = k // 2
    on = mask > 0
    # dilate: dst(y,x) = max_{(i,j) in SE} src(y + i - a, x + j - a)  => a set pixel (y,x) marks
    # dst(y - (i - a), x - (j - a)) for every (i,j) of the element
    dil = np.zeros((h, w), dtype=bool)
    ys, xs = np.nonzero(on)
    for y, x in zip(ys.tolist(), xs.tolist()):
        for i, (j1, j2) in enumerate(rows):
            yy = y - (i - a)
            if j1 >= j2 or yy < 0 or yy >= h:
                continue
            dil[yy, max(x - (j2 - 1) + a, 0):max(min(x - j1 + a + 1, w), 0)] = True
    # erode: dst(y,x) = min_{(i,j) in SE, inside the image} dil(y + i - a, x + j - a)
    pre = np.zeros((h, w + 1), dtype=np.int64)
    pre[:, 1:] = np.cumsum(dil, axis=1)
    out = np.ones((h, w), dtype=bool)
    xs_all = np.arange(w)
    for i, (j1, j2) in enumerate(rows):
        if j1 >= j2:
            continue
        for y in range(h):
            yy = y + i - a
            if yy < 0 or yy >= h:
                continue
            lo = np.clip(xs_all + j1 - a, 0, w)
            hi = np.clip(xs_all + j2 - a, 0, w)
            out[y] &= (pre[yy, hi] - pre[yy, lo]) == (hi - lo)
    return np.where(out, 255, 0).astype(np.uint8)


def point_stamp_ref(layer_rgba: np.ndarray, track: np.ndarray, vis: np.ndarray):
    """STOM.warp_point :163-203.  Returns None (frame untouched) or (cx, cy, radius, rgba)."""
    vis = np.asarray(vis).astype(bool)
    if vis.sum() < len(track) // 2:
        return None
    h, w = layer_rgba.shape[:2]
    alpha = layer_rgba[:, :, 3] > 0
    rgba = layer_rgba[alpha][0].astype(np.int64).tolist() if alpha.any() else [0, 0, 0, 0]
    rgba[3] = max(min(rgba[3], 148), 96)
    mask = np.zeros((h, w), dtype=np.uint8)
    for i in range(len(track)):
        if vis[i]:
            px, py = float(track[i, 1]), float(track[i, 0])
            if not (math.isfinite(px) and math.isfinite(py)):
                return None          # int(nan/inf) raises; the caller's bare except keeps the frame (:93-100)
            x, y = int(px), int(py)  # x is the ROW here, as in the reference
            if 0 <= x < h and 0 <= y < w:
                mask[x, y] = 255
    closed = close_ref(mask, min(h, w) // 15)
    ys, xs = np.nonzero(closed)
    if len(ys) == 0:
        return "empty"              # m00 == 0: an all-transparent layer is composited (frame unchanged)
    m00 = 255.0 * len(ys)
    cx = int((255.0 * float(xs.sum())) / m00)
    cy = int((255.0 * float(ys.sum())) / m00)
    return cx, cy, min(h, w) // 20, tuple(rgba)


def propagate_ref(frames: np.ndarray, layer_rgba: np.ndarray, tracks: np.ndarray, vis: np.ndarray, key_idx: int,
                  shape: str) -> np.ndarray:
    """STOM.propagate_in_video :72-141 on tracker outputs ``tracks [T,N,2]`` (x, y), ``vis [T,N]``."""
    out = []
    h, w = layer_rgba.shape[:2]
    for idx in range(frames.shape[0]):
        frame = frames[idx]
        if idx == key_idx:
            out.append(ov.alpha_composite_ref(frame, layer_rgba))
            continue
        if shape in ("mask", "mask contour"):
            st = point_stamp_ref(layer_rgba, tracks[idx], vis[idx])
            if st is None or st == "empty":
                out.append(frame.copy())
            else:
                cx, cy, r, rgba = st
                out.append(ov.alpha_composite_ref(frame, ov.circle_layer_ref(h, w, cx, cy, r, rgba)))
            continue
        fl = flow_shift_ref(tracks[key_idx], tracks[idx], vis[idx])
        if fl is None:
            out.append(frame.copy())
            continue
        # reference: warp(vip, frame, avg_flow_y (= x flow), avg_flow_x (= y flow)); int64 + float32 -> float64
        out.append(ov.alpha_composite_ref(frame, ov.warp_layer_ref(layer_rgba, float(fl[0]), float(fl[1]))))
    return np.stack(out)
