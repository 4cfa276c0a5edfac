"""Host feed of the visual path: pinned uint8 clips in, merged embeddings out, copies overlapped with compute.

The reference uploads the processor's fp32 ``[M,1176]`` matrix per clip (38.5 MB at 16x448^2, cast to bf16 after the
copy, /root/reference/app.py:434-435); here the decoder's uint8 frames (9.6 MB) are what crosses PCIe, and overlay,
normalisation and patchify run on the GPU.  ``ClipPipeline`` keeps ``depth`` clips in flight: host->device copy of
clip i+1 and device->host copy of result i-1 run on their own streams while clip i is in the tower
(SURVEY.md section 8f rank 4).
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Deque, Optional, Tuple

import torch

from .overlay import OverlaySpec


class ClipPipeline:
    """Double-(or deeper-)buffered ``forward_frames``.

    >>> pipe = ClipPipeline(tower, frames_shape=(16, 448, 448, 3))
    >>> for clip, overlay in clips:                 # clip: uint8 [T,H,W,3] host tensor (pinned for async copies)
    ...     done = pipe.submit(clip, overlay)       # returns the oldest finished result once the pipe is full
    >>> rest = pipe.drain()

    Results are pinned host tensors ``[M/4, out_hidden]``; a returned tensor is valid until the NEXT ``submit`` (the
    result ring has ``depth + 1`` slots, so the slot just handed out is never the one the new clip writes).
    ``after_forward(out_device)`` (optional) is called on the compute stream right after the tower, e.g. to enqueue a
    collective on the device result.
    """

    def __init__(self, tower, frames_shape: Tuple[int, int, int, int], depth: int = 2, to_host: bool = True,
                 after_forward: Optional[Callable[[torch.Tensor], None]] = None):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one clip in the tower, one being copied)")
        t, h, w, c = frames_shape
        if c != 3:
            raise ValueError("frames_shape must be (T, H, W, 3)")
        self.tower, self.depth, self.to_host, self.after_forward = tower, depth, to_host, after_forward
        dev = tower.device
        grid = [[(t + tower.temporal_patch_size - 1) // tower.temporal_patch_size, h // tower.patch_size,
                 w // tower.patch_size]]
        self.grid = grid
        m = grid[0][0] * grid[0][1] * grid[0][2]
        out_dtype = torch.float32 if tower.output_fp32 else tower.dtype
        self.frames_dev = [torch.empty(frames_shape, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.ring = depth + 1                # result slots: one more than clips in flight (see class docstring)
        self.out_dev = [torch.empty((m // tower.spatial_merge_unit, tower.out_hidden_size), dtype=out_dtype, device=dev)
                        for _ in range(self.ring)]
        self.out_host = ([torch.empty(self.out_dev[0].shape, dtype=out_dtype).pin_memory() for _ in range(self.ring)]
                         if to_host else None)
        self.s_h2d, self.s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.ev_input = [None] * depth       # the tower has consumed frames_dev[b]
        self.ev_compute = [None] * self.ring  # out_dev[r] has been produced
        self.ev_d2h = [None] * self.ring      # out_host[r] is complete
        self.i = 0
        self.inflight: Deque[int] = deque()
        self.h2d_bytes = self.frames_dev[0].numel()
        self.d2h_bytes = self.out_dev[0].numel() * self.out_dev[0].element_size() if to_host else 0

    def submit(self, frames_host: torch.Tensor, overlay: Optional[OverlaySpec] = None):
        """Enqueue one clip.  Returns the oldest result (host tensor, synchronised) once ``depth`` clips are in flight,
        else None.  Nothing here blocks the host except that final wait on a result that is ``depth`` clips old."""
        b = self.i % self.depth               # input slot
        r = self.i % self.ring                # result slot
        self.i += 1
        ready = None
        if len(self.inflight) == self.depth:
            ready = self._collect(self.inflight.popleft())
        cur = torch.cuda.current_stream(self.tower.device)
        with torch.cuda.stream(self.s_h2d):
            if self.ev_input[b] is not None:
                self.s_h2d.wait_event(self.ev_input[b])        # the tower has consumed this input buffer
            self.frames_dev[b].copy_(frames_host, non_blocking=True)
            ev_h = torch.cuda.Event()
            ev_h.record(self.s_h2d)
        cur.wait_event(ev_h)
        if self.ev_d2h[r] is not None:
            cur.wait_event(self.ev_d2h[r])                     # the previous result in this slot has left the device
        self.tower.forward_frames(self.frames_dev[b], overlay, grid_thw=self.grid, out=self.out_dev[r])
        if self.after_forward is not None:
            self.after_forward(self.out_dev[r])
        ev = torch.cuda.Event()
        ev.record(cur)
        self.ev_input[b] = ev
        self.ev_compute[r] = ev
        if self.to_host:
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev)
                self.out_host[r].copy_(self.out_dev[r], non_blocking=True)
                self.ev_d2h[r] = torch.cuda.Event()
                self.ev_d2h[r].record(self.s_d2h)
        self.inflight.append(r)
        return ready

    def _collect(self, r: int):
        if self.to_host:
            self.ev_d2h[r].synchronize()
            return self.out_host[r]
        self.ev_compute[r].synchronize()
        return self.out_dev[r]

    def drain(self):
        """Wait for everything in flight; returns the remaining results, oldest first."""
        out = [self._collect(r) for r in self.inflight]
        self.inflight.clear()
        return out

    def fence(self, stream=None):
        """Make ``stream`` (default: current) wait for every enqueued copy, without blocking the host."""
        st = stream if stream is not None else torch.cuda.current_stream(self.tower.device)
        for e in self.ev_d2h + self.ev_compute + self.ev_input:
            if e is not None:
                st.wait_event(e)
