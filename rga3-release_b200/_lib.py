"""ctypes binding of libb200vit.so (the C ABI in include/b200vit.h).

There is no CPU fallback: if the shared library is missing this module raises,
and every entry point that needs the GPU returns B200VIT_EARCH / ECUDA which is
turned into a Python exception here.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200VIT_LIB") or os.path.join(HERE, "libb200vit.so")  # B200VIT_LIB: experiment builds only

# epilogues / enums (keep in sync with include/b200vit.h)
(EPI_STORE_F32, EPI_QKV_ROPE, EPI_BIAS_RESIDUAL, EPI_SWIGLU, EPI_BIAS_GELU, EPI_BIAS_BF16, EPI_BIAS_F32,
 EPI_BIAS_RESIDUAL_NORM, EPI_QKV_ROPE_WINATTN) = range(9)
GEMM_SYNC_INTS = 4096
VERSION = 3
LAYER_NONE, LAYER_RGBA, LAYER_PALETTE, LAYER_BOX = range(4)
FRAME_NONE, FRAME_LAYER, FRAME_CIRCLE = range(3)
(PLAN_M, PLAN_WINDOW_INDEX, PLAN_REVERSE_INDEX, PLAN_CU_WINDOW, PLAN_CU_FULL, PLAN_ROW_MAP, PLAN_ROPE_COS,
 PLAN_ROPE_SIN, PLAN_POS_IDS, PLAN_ROPE_TABLE, PLAN_ROPE_POS) = range(11)


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("depth", "hidden", "intermediate", "heads", "out_hidden", "patch",
                                         "temporal_patch", "merge", "window", "in_channels", "n_fullatt")] + \
               [("fullatt", C.c_int32 * 64)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("qkv_w", "qkv_b", "proj_w", "proj_b", "gateup_w", "gateup_b", "down_w", "down_b")]


class RawLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("norm1_w", "qkv_w", "qkv_b", "proj_w", "proj_b", "norm2_w", "gate_w", "gate_b",
                                          "up_w", "up_b", "down_w", "down_b")]


class RawWeights(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("patch_w", C.c_void_p), ("layers", C.POINTER(RawLayer)),
                ("merger_ln_w", C.c_void_p), ("merger_fc1_w", C.c_void_p), ("merger_fc1_b", C.c_void_p),
                ("merger_fc2_w", C.c_void_p), ("merger_fc2_b", C.c_void_p)]


class Weights(C.Structure):
    _fields_ = [("patch_w", C.c_void_p), ("layers", C.POINTER(LayerWeights)), ("merger_ln_w", C.c_void_p),
                ("merger_fc1_w", C.c_void_p), ("merger_fc1_b", C.c_void_p), ("merger_fc2_w", C.c_void_p),
                ("merger_fc2_b", C.c_void_p), ("ipad", C.c_int32)]


class FrameOp(C.Structure):
    _fields_ = [("mode", C.c_int32), ("sx", C.c_int32), ("sy", C.c_int32), ("zx", C.c_int32), ("zy", C.c_int32),
                ("cx", C.c_int32), ("cy", C.c_int32), ("r", C.c_int32), ("rgba", C.c_uint8 * 4)]


class Overlay(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d_layer", C.c_void_p), ("palette", (C.c_uint8 * 4) * 256),
                ("box", C.c_int32 * 4), ("box_width", C.c_int32), ("h_ops", C.POINTER(FrameOp)),
                ("d_ops", C.c_void_p), ("d_ops_circle_r", C.c_int32)]


class Frames(C.Structure):
    _fields_ = [("d_frames", C.c_void_p), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [("d_a", C.c_void_p), ("d_b", C.c_void_p), ("d_out", C.c_void_p), ("d_bias", C.c_void_p),
                ("d_row_map", C.c_void_p), ("d_rope", C.c_void_p), ("d_rope_pos", C.c_void_p), ("m", C.c_int32),
                ("n", C.c_int32), ("k", C.c_int32), ("ldo", C.c_int32),
                ("epilogue", C.c_int32), ("d_out_bf16", C.c_void_p), ("d_rowsq_out", C.c_void_p),
                ("d_rowsq_in", C.c_void_p), ("rowsq_parts", C.c_int32), ("norm_eps", C.c_float), ("d_sync", C.c_void_p)]


EXPORTS = {
    "b200vit_version": (C.c_int, []),
    "b200vit_last_error": (C.c_char_p, []),
    "b200vit_plan_create": (C.c_int, [C.POINTER(C.c_int64), C.c_int, C.POINTER(Cfg), C.POINTER(C.c_void_p)]),
    "b200vit_plan_destroy": (None, [C.c_void_p]),
    "b200vit_plan_get": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "b200vit_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "b200vit_forward": (C.c_int, [C.c_void_p, C.POINTER(Weights), C.c_void_p, C.POINTER(Frames), C.POINTER(Overlay),
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200vit_forward_launches": (C.c_int, [C.c_void_p, C.c_int]),
    "b200vit_set_l2_persist": (C.c_int, [C.c_int]),
    "b200vit_packed_weights_bytes": (C.c_size_t, [C.POINTER(Cfg)]),
    "b200vit_pack_weights": (C.c_int, [C.POINTER(Cfg), C.POINTER(RawWeights), C.c_void_p, C.c_size_t, C.POINTER(Weights),
                                       C.POINTER(LayerWeights), C.c_void_p]),
    "b200vit_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "b200vit_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "b200vit_overlay_composite": (C.c_int, [C.POINTER(Frames), C.POINTER(Overlay), C.c_void_p, C.c_void_p]),
    "b200vit_overlay_patchify": (C.c_int, [C.POINTER(Frames), C.POINTER(Overlay), C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p]),
    "b200vit_clock_probe": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "b200vit_resize_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "b200vit_resize_bicubic": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200vit_stom_policy_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "b200vit_stom_policy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200vit_raster_polygons": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32,
                                          C.c_uint8, C.c_void_p, C.c_void_p]),
    "b200vit_raster_lines": (C.c_int, [C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint8,
                                       C.c_void_p, C.c_void_p]),
    "b200vit_scribble_points": (C.c_int, [C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double)]),
    "b200vit_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "b200vit_rmsnorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "b200vit_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_void_p]),
    "b200vit_cast_to_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
}

KERNEL_KINDS = ["overlay_patchify", "patch_embed", "rmsnorm", "qkv_rope", "attn_window", "attn_full", "proj_resid",
                "gateup_swiglu", "down_resid", "merger_fc1", "merger_fc2", "qkv_rope_winattn"]

_lib = None


class B200VitError(RuntimeError):
    pass


def lib():
    """Load libb200vit.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200VitError(
                f"{LIB_PATH} not found: build it with `python rga3-release_b200/build.py` "
                "(there is no CPU or PyTorch fallback for this path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().b200vit_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise B200VitError(f"{what}: {msg} (code {rc})")
