"""Same-box comparator: the HF Qwen2.5-VL vision tower in bf16 on the B200, as the
reference runs it (attn_implementation flash_attention_2 / sdpa), on the cfg-2 clip (pixel_values resident).
Not a bench.py line -- context for how the B200-native path compares with the stock GPU path."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hf_ref

DEV = "cuda"
grid = torch.tensor([[8, 32, 32]], device=DEV)
x = torch.randn(8192, 1176, device=DEV).to(torch.bfloat16)
res = {}
for attn in ("flash_attention_2", "sdpa"):
    try:
        model, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.bfloat16, attn=attn, device=DEV)
        with torch.no_grad():
            for _ in range(3):
                hf_ref.hf_forward(model, x, grid)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                hf_ref.hf_forward(model, x, grid)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        ms = sorted(ts)[len(ts) // 2]
        res[attn] = {"ms_per_clip": ms, "frames_per_s": 16 / (ms * 1e-3)}
        print(attn, res[attn], flush=True)
        del model
        torch.cuda.empty_cache()
    except Exception as ex:  # flash-attn may be unusable on this box
        res[attn] = {"error": repr(ex)[:200]}
        print(attn, "failed:", repr(ex)[:200], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/hf_gpu_tower.json", "w"), indent=1)
