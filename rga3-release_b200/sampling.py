"""Frame sampling and the decode hand-off in front of the visual path.

Mirrors, index for index:
  /root/reference/utils/utils.py:201-229           uniform_sample, get_sparse_indices, get_dense_indices
  /root/reference/utils/video_capture.py:10-60     VideoCapture.load_frames_from_video (index choice, BGR->RGB, padding)
  /root/reference/evaluation/videoinfer/inference_videoinfer.py:77-79   sparse indices + the prompted key frame
The decoder itself (cv2 / JPEG files) stays outside: ``stage_clip`` takes whatever it produced -- a list of HxWx3 uint8
arrays or one [N,H,W,3] array -- and fills a PINNED [T,H,W,3] buffer in sampling order, so the uint8 frames cross PCIe
once, asynchronously (``ClipPipeline.submit``), instead of the reference's fp32 [M,1176] upload.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def uniform_sample(total_len: int, sample_num: int) -> List[int]:
    """Middle frame of each of ``sample_num`` equal intervals of [0, total_len) (utils.py:201-208)."""
    intervals = np.linspace(start=0, stop=total_len, num=sample_num + 1).astype(int)
    return [int((intervals[i] + intervals[i + 1] - 1) // 2) for i in range(sample_num)]


def get_sparse_indices(total_frame_num: int, num_frames_mllm: int) -> List[int]:
    """utils.py:211-220: long videos are sampled uniformly; short ones repeat every frame ``num // total`` times and
    add a uniform sample of the remainder."""
    if total_frame_num > num_frames_mllm:
        return sorted(uniform_sample(total_frame_num, num_frames_mllm))
    num_repeat = num_frames_mllm // total_frame_num
    num_sample = num_frames_mllm % total_frame_num
    return sorted(list(range(total_frame_num)) * num_repeat + uniform_sample(total_frame_num, num_sample))


def get_dense_indices(num_frames_mllm: int, num_frames_sam: int) -> List[int]:
    """utils.py:223-229 (note the ``stop = num_frames_mllm - 1``)."""
    intervals = np.linspace(start=0, stop=num_frames_mllm - 1, num=num_frames_sam + 1).astype(int)
    return [int((intervals[i] + intervals[i + 1] - 1) // 2) for i in range(num_frames_sam)]


def video_frame_indices(vlen: int, num_frames: int, sample: str = "uniform", rng: Optional[random.Random] = None) -> List[int]:
    """video_capture.py:25-39: ``min(num_frames, vlen)`` intervals; the middle frame ('uniform') or a random frame of
    each interval ('rand', ``random.choice(range(lo, hi))`` -- the interval's last frame is never drawn)."""
    acc = min(num_frames, vlen)
    intervals = np.linspace(start=0, stop=vlen, num=acc + 1).astype(int)
    ranges = [(int(intervals[i]), int(intervals[i + 1]) - 1) for i in range(acc)]
    if sample == "rand":
        r = rng or random
        return [r.choice(range(lo, hi)) for lo, hi in ranges]
    return [(lo + hi) // 2 for lo, hi in ranges]


def clip_indices_with_key_frame(total_frames: int, num_frames: int, key_frame_idx: int) -> Tuple[List[int], int]:
    """inference_videoinfer.py:77-79: ``num_frames - 1`` sparse indices plus the frame the user prompted, sorted.
    Returns (indices, position of the key frame inside them) -- the ``key_idx`` STOM.propagate_in_video is given."""
    idxs = sorted(get_sparse_indices(total_frames, num_frames - 1) + [int(key_frame_idx)])
    return idxs, idxs.index(int(key_frame_idx))


def stage_clip(frames, indices: Sequence[int], num_frames: Optional[int] = None, bgr: bool = False,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decoded frames -> one pinned uint8 [T,H,W,3] clip in sampling order.
    ``frames``: sequence of HxWx3 uint8 arrays (what ``cap.read()`` / PIL decode yield) or an [N,H,W,3] array / tensor.
    ``bgr=True`` flips the channel order (cv2.cvtColor(frame, COLOR_BGR2RGB), video_capture.py:52).  When fewer than
    ``num_frames`` indices are given the last frame is repeated (video_capture.py:58-59).  ``out``: reuse a buffer."""
    t = int(num_frames) if num_frames is not None else len(indices)
    if len(indices) == 0 or t < len(indices):
        raise ValueError("stage_clip needs at least one index and num_frames >= len(indices)")
    first = np.asarray(frames[int(indices[0])])
    if first.ndim != 3 or first.shape[2] != 3 or first.dtype != np.uint8:
        raise ValueError("frames must be HxWx3 uint8")
    h, w = first.shape[:2]
    if out is None:
        out = torch.empty((t, h, w, 3), dtype=torch.uint8)
        if torch.cuda.is_available():
            out = out.pin_memory()
    elif tuple(out.shape) != (t, h, w, 3) or out.dtype != torch.uint8 or out.is_cuda:
        raise ValueError(f"out must be a host uint8 tensor of shape {(t, h, w, 3)}")
    dst = out.numpy()
    for k in range(t):
        src = np.asarray(frames[int(indices[min(k, len(indices) - 1)])])
        if src.shape != (h, w, 3):
            raise ValueError("all frames of a clip must have one size")
        dst[k] = src[:, :, ::-1] if bgr else src
    return out
