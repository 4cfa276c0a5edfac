"""Run the attention kernel a few times (for ncu): python tools/prof_attn.py <seg_len> [iters]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_ops
from rga3_release_b200 import _lib

seg = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
M, D = bench_ops.M, 1280
qkv = torch.randn(M, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(M, D, dtype=torch.bfloat16, device="cuda")
cu = np.arange(0, M + 1, seg, dtype=np.int32)


def fn():
    _lib.check(_lib.lib().b200vit_attention(qkv.data_ptr(), out.data_ptr(), cu.ctypes.data_as(C.POINTER(C.c_int32)),
                                            len(cu) - 1, 16, bench_ops.stream()), "attn")


for _ in range(iters):
    fn()
torch.cuda.synchronize()
med, mn = bench_ops.timeit(fn, iters=10)
print("attn seg", seg, "median us", med * 1e3)
