"""Run one GEMM shape a few times (for ncu): python tools/prof_gemm.py <name> [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_ops
from rga3_release_b200 import _lib

M, D, IPAD = bench_ops.M, 1280, 3456
SHAPES = {
    "patch_embed": (M, D, 1176, _lib.EPI_STORE_F32, None, False, torch.float32),
    "qkv": (M, 3 * D, D, _lib.EPI_QKV_ROPE, None, True, torch.bfloat16),
    "qkvwin": (M, 3 * D, D, _lib.EPI_QKV_ROPE_WINATTN, D, True, torch.bfloat16),
    "proj": (M, D, D, _lib.EPI_BIAS_RESIDUAL_NORM, None, False, torch.float32),
    "gateup": (M, 2 * IPAD, D, _lib.EPI_SWIGLU, IPAD, False, torch.bfloat16),
    "down": (M, D, IPAD, _lib.EPI_BIAS_RESIDUAL_NORM, None, False, torch.float32),
}
name = sys.argv[1] if len(sys.argv) > 1 else "gateup"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
fn = bench_ops.gemm_fn(*SHAPES[name])
for _ in range(iters):
    fn()
torch.cuda.synchronize()
med, mn = bench_ops.timeit(fn, iters=10)
print(name, "median us", med * 1e3)
