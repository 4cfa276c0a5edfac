"""Secondary numbers for the other BASELINE.json configs on ONE B200 (bench.py is the judged cfg-2 line):
   cfg1/2 [8,32,32], cfg3 clip [16,32,32] (8 of them = cfg 3 on one GPU), cfg4 [32,48,48] (64-frame 672x672)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import rga3_release_b200 as vit

DEV = "cuda"
tower = vit.B200VisionTower(dict(bench.CFG_7B), device=DEV, return_dict=False)
bench.random_state_dict_gpu(tower)
peaks = bench.measured_peaks()
res = {}
for name, (t, h, w), reps in (("cfg2_16x448", (16, 448, 448), 20), ("cfg3_clip_32x448", (32, 448, 448), 10), ("cfg4_64x672", (64, 672, 672), 5)):
    g = torch.Generator(device=DEV).manual_seed(1)
    frames = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, device=DEV, generator=g)
    grid = [[t // 2, h // 14, w // 14]]
    for _ in range(3):
        tower.forward_frames(frames)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        tower.forward_frames(frames)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    flops, _ = bench.algorithmic_flops(grid)
    res[name] = {"grid_thw": grid, "patches": grid[0][0] * grid[0][1] * grid[0][2], "ms_per_clip": ms, "frames_per_s": t / (ms * 1e-3),
                 "merged_tokens_per_s": grid[0][0] * grid[0][1] * grid[0][2] / 4 / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12,
                 "frac_bf16_burst_peak": flops / (ms * 1e-3) / 1e12 / peaks["bf16_burst"],
                 "workspace_gb": tower.plan_for(grid).ws_bytes / 1e9}
    tower.profile(grid, True)
    tower.forward_frames(frames)
    prof = tower.profile_read(grid)
    tower.profile(grid, False)
    _, per = bench.algorithmic_flops(grid)
    res[name]["kernel_ms"] = {k: round(v[0], 3) for k, v in prof.items() if v[1]}
    res[name]["kernel_tflops"] = {k: round(per[k] / (v[0] * 1e-3) / 1e12, 1) for k, v in prof.items() if v[1] and k in per}
    print(name, json.dumps(res[name]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_configs.json", "w"), indent=1)
