"""Host side of the STOM visual-prompt overlay: the prompt layer description and
the per-frame placement policy.  The pixel work (translate, stamp, composite,
normalise, patchify) runs in csrc/overlay.cu.

Mirrors /root/reference/model/STOM.py:
  propagate_in_video :72-141  -> ``stom_frame_ops``  (policy: which frame gets what)
  warp              :145-160  -> FrameOp(mode=LAYER, sx, sy, zx, zy)
  warp_point        :163-207  -> FrameOp(mode=CIRCLE, cx, cy, r, rgba)
and the layer construction of /root/reference/utils/visual_prompt_generator.py
(image_blending :284-368: box :102-104, mask :268-274, scribble :230-252).
CoTracker itself (STOM.track_in_video :25-69) is a separate third-party model and
stays outside; its outputs (tracks, visibility) are the inputs of ``stom_frame_ops``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


@dataclass
class FrameOp:
    mode: int = _lib.FRAME_NONE
    sx: int = 0
    sy: int = 0
    zx: int = 0
    zy: int = 0
    cx: int = 0
    cy: int = 0
    r: int = 0
    rgba: Tuple[int, int, int, int] = (0, 0, 0, 0)


def shift_from_flow(flow: float, n: int) -> Tuple[int, int]:
    """Integer form of STOM.warp's ``int(x + flow)`` (STOM.py:151-152) along one
    axis of length n: (shift, zero_extra).  Python's int() truncates toward
    zero, so a source that lands in (-1, 0) is written to coordinate 0 as well;
    zero_extra marks that extra (earlier) source for destination 0."""
    coords = np.arange(n, dtype=np.int64) + flow            # float64, as in the reference
    dst = np.trunc(coords).astype(np.int64)
    nonneg = coords >= 0
    if nonneg.any():
        d = np.unique(dst[nonneg] - np.arange(n, dtype=np.int64)[nonneg])
        if d.size != 1:
            raise ValueError("flow does not reduce to an integer shift")
        shift = int(d[0])
    else:
        shift = int(np.floor(flow))
    zero_extra = int(np.any((~nonneg) & (dst == 0)))
    return shift, zero_extra


FRAME_OP_BYTES = C.sizeof(_lib.FrameOp)  # 36


@dataclass
class OverlaySpec:
    """Prompt layer + per-frame ops.  ``layer`` is a CUDA uint8 tensor:
    [H,W,4] for kind RGBA, [H,W] palette indices (0 = transparent) for PALETTE."""
    kind: int = _lib.LAYER_NONE
    layer: Optional[torch.Tensor] = None
    palette: np.ndarray = field(default_factory=lambda: np.zeros((256, 4), dtype=np.uint8))
    box: Tuple[int, int, int, int] = (0, 0, 0, 0)
    box_width: int = 1
    ops: List[FrameOp] = field(default_factory=list)
    # device-resident ops (``stom_frame_ops_device``): uint8 CUDA tensor [T, 36] holding b200vit_frame_op records;
    # used instead of ``ops`` when set, so nothing about the placement ever visits the host
    device_ops: Optional[torch.Tensor] = None
    device_ops_circle_r: int = -1

    # ---- constructors mirroring image_blending's shapes
    @classmethod
    def from_rgba(cls, layer_rgba, ops: Sequence[FrameOp], device="cuda"):
        lay = torch.as_tensor(np.ascontiguousarray(layer_rgba) if isinstance(layer_rgba, np.ndarray) else layer_rgba)
        assert lay.dtype == torch.uint8 and lay.dim() == 3 and lay.shape[2] == 4
        return cls(kind=_lib.LAYER_RGBA, layer=lay.to(device).contiguous(), ops=list(ops))

    @classmethod
    def from_palette(cls, index_layer, palette, ops: Sequence[FrameOp], device="cuda"):
        """1 byte/pixel coverage: index k > 0 shows palette[k] (rgba); covers mask and scribble
        prompts (PIL rasterises them on the host, visual_prompt_generator.py:230-274)."""
        idx = torch.as_tensor(np.ascontiguousarray(index_layer) if isinstance(index_layer, np.ndarray) else index_layer)
        assert idx.dtype == torch.uint8 and idx.dim() == 2
        pal = np.zeros((256, 4), dtype=np.uint8)
        p = np.asarray(palette, dtype=np.uint8).reshape(-1, 4)
        pal[: p.shape[0]] = p
        return cls(kind=_lib.LAYER_PALETTE, layer=idx.to(device).contiguous(), palette=pal, ops=list(ops))

    @classmethod
    def from_box(cls, box, width: int, rgba, ops: Sequence[FrameOp]):
        """Analytic rectangle outline, 0 bytes/pixel (draw_rectangle, visual_prompt_generator.py:102-104)."""
        pal = np.zeros((256, 4), dtype=np.uint8)
        pal[1] = np.asarray(rgba, dtype=np.uint8)
        return cls(kind=_lib.LAYER_BOX, palette=pal, box=tuple(int(v) for v in box), box_width=max(int(width), 1),
                   ops=list(ops))

    def validate(self, h: int, w: int, device):
        """The prompt layer must cover the frames exactly: the kernel indexes it with the frame's pitch, and the
        reference's ``Image.alpha_composite`` (STOM.py:84-87, :157-160) raises on a size mismatch as well."""
        if self.kind in (_lib.LAYER_RGBA, _lib.LAYER_PALETTE):
            lay = self.layer
            want = (h, w, 4) if self.kind == _lib.LAYER_RGBA else (h, w)
            if lay is None or tuple(lay.shape) != want or lay.dtype != torch.uint8 or not lay.is_contiguous():
                raise ValueError(f"overlay layer must be a contiguous uint8 tensor of shape {want} matching the frames, "
                                 f"got {None if lay is None else tuple(lay.shape)} (resize the prompt layer with the frames)")
            if not lay.is_cuda or lay.device != torch.device(device):
                raise ValueError(f"overlay layer is on {lay.device}, frames on {device}")
        if self.device_ops is not None and self.device_ops.device != torch.device(device):
            raise ValueError(f"overlay device_ops are on {self.device_ops.device}, frames on {device}")

    # ---- C view (keeps the backing arrays alive on self)
    def to_c(self, n_frames: int):
        ov = _lib.Overlay()
        ov.kind = self.kind
        ov.d_layer = self.layer.data_ptr() if self.layer is not None else None
        C.memmove(ov.palette, np.ascontiguousarray(self.palette, dtype=np.uint8).ctypes.data, 1024)
        for i in range(4):
            ov.box[i] = int(self.box[i])
        ov.box_width = int(self.box_width)
        if self.device_ops is not None:
            d = self.device_ops
            if not (d.is_cuda and d.dtype == torch.uint8 and d.is_contiguous() and d.numel() >= n_frames * FRAME_OP_BYTES):
                raise ValueError("device_ops must be a contiguous CUDA uint8 tensor of at least [T, 36] bytes")
            ov.h_ops = None
            ov.d_ops = d.data_ptr()
            ov.d_ops_circle_r = int(self.device_ops_circle_r)
            ov._refs = (d, self.layer)
            self._c_keep = (ov, d)
            return ov
        ops = (_lib.FrameOp * n_frames)()
        for i in range(n_frames):
            o = self.ops[i] if i < len(self.ops) else FrameOp()
            ops[i].mode, ops[i].sx, ops[i].sy, ops[i].zx, ops[i].zy = o.mode, o.sx, o.sy, o.zx, o.zy
            ops[i].cx, ops[i].cy, ops[i].r = o.cx, o.cy, o.r
            for j in range(4):
                ops[i].rgba[j] = int(o.rgba[j])
        ov.h_ops = ops
        ov._refs = (ops, self.layer)   # the C view keeps its backing storage alive, even if this spec is a temporary
        self._c_keep = (ov, ops)
        return ov


def stom_frame_ops(pred_tracks: np.ndarray, pred_visibility: np.ndarray, key_idx: int, shape: str, h: int, w: int,
                   layer_rgba: np.ndarray) -> List[FrameOp]:
    """The placement policy of STOM.propagate_in_video (STOM.py:72-141) on tracker
    outputs ``pred_tracks [T,N,2] (x,y)`` / ``pred_visibility [T,N]``: key frame
    gets the layer unmoved; mask shapes get a circle stamp at the centroid of the
    visible tracks (warp_point); other shapes the layer translated by the mean
    of the MAD-filtered flows (warp); frames the reference leaves untouched get
    FRAME_NONE."""
    import cv2  # the centroid uses OpenCV's closing + moments exactly as the reference does (:187-193)
    t_frames = pred_tracks.shape[0]
    ops: List[FrameOp] = []
    key_track = pred_tracks[key_idx]
    alpha_mask = layer_rgba[:, :, 3] > 0
    for idx in range(t_frames):
        if idx == key_idx:
            ops.append(FrameOp(mode=_lib.FRAME_LAYER))
            continue
        trk, vis = pred_tracks[idx], pred_visibility[idx].astype(bool)
        if shape in ("mask", "mask contour"):
            if vis.sum() < len(trk) // 2:                                  # :165-166
                ops.append(FrameOp())
                continue
            if alpha_mask.any():
                rgba = layer_rgba[alpha_mask][0].astype(np.int64).tolist()  # :169-170
            else:
                rgba = [0, 0, 0, 0]
            rgba[3] = max(min(rgba[3], 148), 96)                           # :174
            if not np.isfinite(trk[vis]).all():
                # int(nan) / int(inf) raises in warp_point (:182-183); the caller's bare except keeps the frame (:93-100)
                ops.append(FrameOp())
                continue
            mask = np.zeros((h, w), dtype=np.uint8)
            for i, pt in enumerate(trk):                                   # :180-185
                if vis[i]:
                    xx, yy = int(pt[1].item()), int(pt[0].item())
                    if 0 <= xx < h and 0 <= yy < w:
                        mask[xx, yy] = 255
            ks = min(h, w) // 15
            kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ks, ks))
            closed = cv2.morphologyEx(mask, cv2.MORPH_CLOSE, kernel)
            mom = cv2.moments(closed)
            if mom["m00"] != 0:
                ops.append(FrameOp(mode=_lib.FRAME_CIRCLE, cx=int(mom["m10"] / mom["m00"]), cy=int(mom["m01"] / mom["m00"]),
                                   r=min(h, w) // 20, rgba=tuple(rgba)))
            else:
                ops.append(FrameOp())
            continue
        flows = trk[vis] - key_track[vis]                                  # :104-106
        if len(flows) == 0:
            ops.append(FrameOp())
            continue
        mag = np.linalg.norm(flows, axis=1)
        med = np.median(mag)
        mad = np.median(np.abs(mag - med))
        keep = (mag >= med - 3 * mad) & (mag <= med + 3 * mad)             # :112-118
        filt = flows[keep]
        if len(filt) < vis.shape[0] // 2:                                  # :122
            ops.append(FrameOp())
            continue
        fx = np.mean(filt[:, 0]) if filt.size > 0 else 0.0                 # the reference's "avg_flow_y" (:126-127)
        fy = np.mean(filt[:, 1]) if filt.size > 0 else 0.0
        if np.isnan(fx) or np.isnan(fy):
            ops.append(FrameOp())
            continue
        sx, zx = shift_from_flow(fx, w)
        sy, zy = shift_from_flow(fy, h)
        ops.append(FrameOp(mode=_lib.FRAME_LAYER, sx=sx, sy=sy, zx=zx, zy=zy))
    return ops


def stom_frame_ops_device(pred_tracks: torch.Tensor, pred_visibility: torch.Tensor, key_idx: int, shape: str, h: int, w: int,
                          layer_rgba: Optional[torch.Tensor] = None, stream=None) -> Tuple[torch.Tensor, int]:
    """``stom_frame_ops`` without the host: the same placement policy (STOM.py:72-141) computed by
    csrc/stom_policy.cu from the tracker's CUDA outputs ``pred_tracks [T,N,2]`` fp32 (x, y) and
    ``pred_visibility [T,N]`` (bool/uint8).  Returns ``(device_ops, circle_r)`` for
    ``OverlaySpec(device_ops=..., device_ops_circle_r=...)``: a uint8 CUDA tensor [T, 36] of
    b200vit_frame_op records, and the shared circle radius (-1 for non-mask shapes).  No device->host copy,
    no synchronisation: the ops are consumed by the overlay kernel on the same stream."""
    if not (pred_tracks.is_cuda and pred_visibility.is_cuda):
        raise ValueError("stom_frame_ops_device takes CUDA tensors (use stom_frame_ops for host arrays)")
    if pred_tracks.dim() != 3 or pred_tracks.shape[2] != 2 or tuple(pred_visibility.shape) != tuple(pred_tracks.shape[:2]):
        raise ValueError("pred_tracks must be [T,N,2] and pred_visibility [T,N]")
    trk = pred_tracks.to(torch.float32).contiguous()
    vis = pred_visibility.to(torch.uint8).contiguous()
    t, n = int(trk.shape[0]), int(trk.shape[1])
    mask_shape = shape in ("mask", "mask contour")
    lay_ptr = None
    if mask_shape:
        if layer_rgba is None or not layer_rgba.is_cuda or layer_rgba.dtype != torch.uint8 or tuple(layer_rgba.shape) != (h, w, 4):
            raise ValueError("mask shapes need the CUDA uint8 RGBA layer [h,w,4]")
        layer_rgba = layer_rgba.contiguous()
        lay_ptr = layer_rgba.data_ptr()
    l = _lib.lib()
    ws_bytes = int(l.b200vit_stom_policy_workspace_bytes(t, n, h, w))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=trk.device)
    ops = torch.empty((t, FRAME_OP_BYTES), dtype=torch.uint8, device=trk.device)
    s = stream if stream is not None else torch.cuda.current_stream(trk.device).cuda_stream
    with torch.cuda.device(trk.device):
        rc = l.b200vit_stom_policy(trk.data_ptr(), vis.data_ptr(), t, n, int(key_idx), int(mask_shape), int(h), int(w),
                                   lay_ptr, ops.data_ptr(), ws.data_ptr(), ws_bytes, s)
    _lib.check(rc, "stom_policy")
    ops._keep = (trk, vis, ws, layer_rgba)  # inputs stay alive until the enqueued kernels have consumed them
    return ops, (min(h, w) // 20 if mask_shape else -1)


def frame_ops_from_bytes(raw: np.ndarray) -> List[FrameOp]:
    """Decode a [T,36] uint8 array of b200vit_frame_op records (e.g. ``device_ops.cpu().numpy()``)."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, FRAME_OP_BYTES)
    out = []
    for rec in raw:
        iv = rec[:32].view(np.int32)
        out.append(FrameOp(mode=int(iv[0]), sx=int(iv[1]), sy=int(iv[2]), zx=int(iv[3]), zy=int(iv[4]), cx=int(iv[5]),
                           cy=int(iv[6]), r=int(iv[7]), rgba=tuple(int(v) for v in rec[32:36])))
    return out
