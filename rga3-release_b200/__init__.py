"""B200-native visual path for RGA3 (STOM overlay -> patchify -> Qwen2.5-VL vision tower).

Python host side of libb200vit.so; see include/b200vit.h, DESIGN.md, INTEGRATION.md.
"""
from . import _lib
from ._lib import B200VitError, lib
from .module import B200VisionTower, install, splice_span
from .dist import gather_tokens, shard_clips, shard_slices
from .preprocess import fit_frames, resize_frames, smart_resize
from .pipeline import ClipPipeline
from .overlay import (FrameOp, OverlaySpec, frame_ops_from_bytes, shift_from_flow, stom_frame_ops,
                      stom_frame_ops_device)
from .mrope import mrope_position_ids
from .sampling import (clip_indices_with_key_frame, get_dense_indices, get_sparse_indices, stage_clip, uniform_sample,
                       video_frame_indices)
from .prompts import (get_bbox_from_mask, lines_layer, mask_layer, prompt_alpha, prompt_line_width, scribble_layer,
                      scribble_points)

__all__ = ["B200VisionTower", "install", "splice_span", "OverlaySpec", "FrameOp", "shift_from_flow", "stom_frame_ops", "stom_frame_ops_device", "frame_ops_from_bytes", "lib", "shard_clips", "shard_slices", "gather_tokens", "smart_resize", "resize_frames", "fit_frames", "ClipPipeline",
           "B200VitError", "mask_layer", "scribble_layer", "lines_layer", "scribble_points", "prompt_alpha", "prompt_line_width",
           "get_bbox_from_mask", "mrope_position_ids", "uniform_sample", "get_sparse_indices", "get_dense_indices", "video_frame_indices",
           "clip_indices_with_key_frame", "stage_clip"]
