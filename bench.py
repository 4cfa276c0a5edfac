#!/usr/bin/env python
"""Benchmark of the RGA3 visual path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg4-split]

A step = one pass of the hot path (STOM overlay -> normalise -> patchify -> 32-layer Qwen2.5-VL-7B-shaped vision tower
-> merged visual embeddings) over one batch of synthetic input.  Workloads (BASELINE.json `configs`):
  cfg2 (default, the configuration the metric is quoted on): one 16-frame 448x448 clip per GPU per step, box + mask
        prompt overlay with a per-frame shift.  Weak scaling: every rank has its own clip.
  cfg3: a FIXED batch of eight 32-frame 448x448 clips per step, sharded by clip over the ranks (8 / N clips each).
        Strong scaling.
  cfg4: one 64-frame 672x672 clip (73,728 patches) per GPU per step.  Weak scaling.
  cfg4-split: ONE such clip per step, split by temporal slices over the ranks (`shard_slices`: no attention segment
        spans slices).  Strong scaling.
For N > 1 (torchrun, one rank per GPU) the merged tokens of every clip go to rank 0 with NCCL inside the timed region.

`value`  : frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through the public feed API (ClipPipeline) with HOST (pinned) uint8 frames in and HOST
           embeddings out, H2D/D2H inside the timed region.
`cpu_baseline` / `--impl reference`: the reference's own CPU implementation of the path (STOM.warp pixel scatter ->
           PIL alpha_composite -> HF Qwen2VLVideoProcessor -> HF tower fp32 eager) on this box's host cores, on the
           whole cfg2 clip.
`gpu_baseline` (N = 1): the stock GPU path the reference runs (HF tower bf16 + flash-attn-2) on the same box, same run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG_7B = dict(depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=3584,
              window_size=112, fullatt_block_indexes=[7, 15, 23, 31])
T_FRAMES, H, W = 16, 448, 448   # cfg2
WORKLOADS = {
    "cfg2": dict(t=16, hw=448, scaling="weak", desc="cfg2: one 16-frame 448x448 clip, STOM box+mask overlay on all 16 frames, "
                 "grid_thw=[8,32,32], 8192 patches -> 2048 merged tokens"),
    "cfg3": dict(t=32, hw=448, scaling="strong", batch=8, desc="cfg3: fixed batch of eight 32-frame 448x448 clips per step "
                 "(grid_thw=[16,32,32] each, 131072 patches), sharded by clip over the GPUs, STOM overlay on all frames"),
    "cfg4": dict(t=64, hw=672, scaling="weak", desc="cfg4: one 64-frame 672x672 clip per GPU, grid_thw=[32,48,48], "
                 "73728 patches -> 18432 merged tokens, STOM overlay on all frames"),
    "cfg4-split": dict(t=64, hw=672, scaling="strong", split=True, desc="cfg4-split: ONE 64-frame 672x672 clip per step "
                       "(grid_thw=[32,48,48]) split by temporal slices over the GPUs, STOM overlay on all frames"),
}


# ----------------------------------------------------------------------------- algorithmic work
def algorithmic_flops(grid_thw, cfg=CFG_7B):
    """2 FLOP/MAC, no padding, softmax/elementwise excluded.  Returns (total, per-kernel dict)."""
    d, i, o, depth = cfg["hidden_size"], cfg["intermediate_size"], cfg["out_hidden_size"], cfg["depth"]
    kpe = 3 * 2 * 14 * 14
    m = sum(t * h * w for t, h, w in grid_thw)
    per = {"patch_embed": 2.0 * m * kpe * d, "qkv_rope": 2.0 * m * d * 3 * d * depth, "proj_resid": 2.0 * m * d * d * depth,
           "gateup_swiglu": 2.0 * m * d * 2 * i * depth, "down_resid": 2.0 * m * i * d * depth,
           "merger_fc1": 2.0 * (m // 4) * (4 * d) ** 2, "merger_fc2": 2.0 * (m // 4) * 4 * d * o}
    n_full = len(cfg["fullatt_block_indexes"])
    full = sum(t * 4.0 * (h * w) ** 2 * d for t, h, w in grid_thw)
    win = 0.0
    for t, h, w in grid_thw:  # window segments: 4x4 merged units = 8x8 patches, ragged at the edges
        lh, lw = h // 2, w // 2
        for wy in range(0, lh, 4):
            for wx in range(0, lw, 4):
                n = min(4, lh - wy) * min(4, lw - wx) * 4
                win += t * 4.0 * n * n * d
    per["attn_full"] = full * n_full
    per["attn_window"] = win * (depth - n_full)
    return sum(per.values()), per


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(bf16_burst=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    hbm_gbs=j["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, taken from the committed ncu --set full
    capture named in profiles/traffic.json (not re-measured per run: ncu cannot run inside a timed benchmark)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    j = json.load(open(p))
    e = j.get(kernel)
    return (e["bytes_per_launch"], e["source"]) if e else (None, None)


# ----------------------------------------------------------------------------- synthetic workload
def synthetic_frames(t, h, w, clip_id=0):
    g = torch.Generator().manual_seed(1000 + clip_id)
    noise = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, generator=g)
    yy = torch.linspace(0, 1, h).view(1, h, 1, 1)
    xx = torch.linspace(0, 1, w).view(1, 1, w, 1)
    tt = torch.linspace(0, 1, max(t, 2))[:t].view(t, 1, 1, 1)
    smooth = (127.5 + 127.5 * torch.sin(6.28318 * (yy * 1.5 + xx * 0.75 + tt))).expand(t, h, w, 3)
    return ((noise.float() + smooth) * 0.5).round().clamp(0, 255).to(torch.uint8)


def prompt_layer(hw=448):
    """cfg 2 prompt: red box outline (112,96,335,351) width 4 alpha 200 + lime disc r=80 alpha 100 at
    448x448; scaled with the frame for 672x672."""
    from PIL import Image, ImageDraw
    s = hw / 448.0
    vip = Image.new("RGBA", (hw, hw), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    c, r = int(224 * s), int(80 * s)
    d.ellipse([(c - r, c - r), (c + r, c + r)], fill=(0, 255, 0, 100))
    d.rectangle([(int(112 * s), int(96 * s)), (int(335 * s), int(351 * s))], outline=(255, 0, 0, 200), width=max(int(4 * s), 1))
    return np.array(vip)


def frame_shifts(t):
    """per-frame integer translation of the prompt (the output of the STOM policy): dx = dy = frame - t/2, clamped"""
    return [max(-24, min(24, i - t // 2)) for i in range(t)]


def random_state_dict_gpu(tower, seed=0):
    """Random-init weights of the 7B tower shape, generated on the device (N(0, 0.02) matrices and
    biases, norm weights 1 + 0.1 N(0,1))."""
    g = torch.Generator(device=tower.device).manual_seed(seed)
    with torch.no_grad():
        for name, p in tower.named_parameters():
            if name.endswith(("norm1.weight", "norm2.weight", "ln_q.weight")):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=p.device))
    tower.invalidate()


def config_dict(workload, world):
    """The `config` both arms print (the reference arm runs the same workload definition on the host cores)."""
    wl = WORKLOADS[workload]
    par = "single GPU" if world == 1 else (
        f"dp{world}: one clip per GPU, merged tokens gathered to rank 0" if wl["scaling"] == "weak" and not wl.get("split") else
        f"dp{world}: the step's clips sharded by clip, merged tokens gathered to rank 0" if not wl.get("split") else
        f"dp{world}: the clip's temporal slices sharded over the GPUs, merged tokens gathered to rank 0")
    return {"workload": wl["desc"], "weights": "random-init Qwen2.5-VL-7B vision tower shape (676.6M params)",
            "parallelism": par,
            "l2": "per-step working set (1.35 GB bf16 weights + >= 0.3 GB activations) exceeds the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- clocks
_SAMPLER_SRC = r"""
import sys, time, pynvml as nv
idx, interval = int(sys.argv[1]), float(sys.argv[2])
calls = sys.argv[3] if len(sys.argv) > 3 else "cpr"
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(idx)
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)   # ~4 ms per call: once, up front
while True:
    r = 0
    if "r" in calls:
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    p = nv.nvmlDeviceGetPowerUsage(h) / 1000.0 if "p" in calls else 0.0
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), p, int(r), flush=True)
    time.sleep(interval)
"""


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region, read through NVML in a SEPARATE process
    (a sampling thread inside the benchmark process slowed multi-GPU steps by 15-25 %: its NVML calls contend
    with the launching thread; measured)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, interval=0.1):
        self.rows, self.proc, self.max_mhz = [], None, None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        self.interval = interval

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.phys), str(self.interval),
                                          os.environ.get("BENCH_SAMPLER_CALLS", "cpr")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            first = self.proc.stdout.readline().split()      # blocks until NVML is up in the child
            self.max_mhz = float(first[1]) if len(first) == 2 and first[0] == "max" else None
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = line.split()
            if len(f) == 4:
                self.rows.append((float(f[0]), float(f[1]), float(f[2]), int(f[3])))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-1:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock sample inside the timed region"]}
        bits = 0
        for r in rows:
            bits |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(rows),
                "power_w_max": float(max(r[2] for r in rows)), "source": "nvml (separate process)"}


# ----------------------------------------------------------------------------- reference (CPU) arm
def reference_step_fn():
    """The reference's CPU path on the whole cfg2 clip: for every frame STOM.warp's per-pixel scatter of the prompt
    layer (model/STOM.py:145-155, restated in oracle/overlay_ref.py) -> PIL alpha_composite (:157-160) -> HF
    Qwen2VLVideoProcessor -> HF tower fp32 eager on all host cores."""
    from PIL import Image
    from oracle import hf_ref, overlay_ref
    torch.set_num_threads(os.cpu_count() or 1)
    model, cfg, _ = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.float32, attn="eager")
    frames = synthetic_frames(T_FRAMES, H, W, 0).numpy()
    layer = prompt_layer(H)
    shifts = frame_shifts(T_FRAMES)
    parts = {}

    def step():
        t0 = time.perf_counter()
        comp = []
        for i in range(T_FRAMES):
            warped = overlay_ref.warp_layer_ref(layer, float(shifts[i]), float(shifts[i]))
            pil = Image.alpha_composite(Image.fromarray(frames[i], "RGB").convert("RGBA"), Image.fromarray(warped, "RGBA"))
            comp.append(np.array(pil.convert("RGB")))
        t1 = time.perf_counter()
        pv, grid = hf_ref.hf_patchify(np.stack(comp))
        t2 = time.perf_counter()
        out = hf_ref.hf_forward(model, pv, grid)
        t3 = time.perf_counter()
        parts.update(overlay_s=t1 - t0, processor_s=t2 - t1, tower_s=t3 - t2)
        return out
    return step, T_FRAMES, parts


def cpu_baseline_entry(fps, parts, timed_calls):
    return {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"the whole cfg2 clip (16 frames, grid_thw=[8,32,32]), {timed_calls} timed call(s) after 1 warm-up: STOM.warp "
                      "pixel scatter + PIL alpha_composite + HF video processor + HF tower fp32 eager on the host cores",
            "seconds_per_clip": {k: round(v, 3) for k, v in parts.items()}}


def run_reference(args, rank, world):
    if rank != 0:
        return
    step, nf, parts = reference_step_fn()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = nf * args.steps / dt
    line = {"impl": "reference", "metric": "vision_tower_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict("cfg2", max(args.gpus, 1)),
            "cpu_baseline": cpu_baseline_entry(fps, parts, args.steps),
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tokens_per_s": fps * 128.0}
    print(json.dumps(line), flush=True)


def gpu_baseline(dev, pv, grid):
    """The stock GPU path of the reference (app.py:50-56: bf16, flash_attention_2) on this box: HF tower on the same
    pixel_values, resident, CUDA-event timed."""
    from oracle import hf_ref
    res = {}
    for attn in ("flash_attention_2", "sdpa"):
        try:
            model, _, _ = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.bfloat16, attn=attn, device=dev)
            with torch.no_grad():
                for _ in range(3):
                    hf_ref.hf_forward(model, pv, grid)
                torch.cuda.synchronize(dev)
                ts = []
                for _ in range(10):
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    hf_ref.hf_forward(model, pv, grid)
                    e.record()
                    torch.cuda.synchronize(dev)
                    ts.append(s.elapsed_time(e))
            ms = float(np.median(ts))
            res[attn] = {"ms_per_clip": ms, "frames_per_s": T_FRAMES / (ms * 1e-3)}
            del model
            torch.cuda.empty_cache()
        except Exception as ex:  # flash-attn may be unusable on a box
            res[attn] = {"error": repr(ex)[:160]}
    best = min((v["ms_per_clip"], k) for k, v in res.items() if "ms_per_clip" in v) if any("ms_per_clip" in v for v in res.values()) else None
    return {"what": "HF Qwen2_5_VisionTransformerPretrainedModel bf16 on this GPU, pixel_values resident (overlay / patchify not "
                    "included), cfg2 clip", "attn": res,
            "value": (T_FRAMES / (best[0] * 1e-3)) if best else None, "unit": "frames/s", "best_attn": best[1] if best else None}


# ----------------------------------------------------------------------------- ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BENCH_WORKLOAD", "cfg2"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import rga3_release_b200 as vit
    from rga3_release_b200 import _lib

    wl = WORKLOADS[args.workload]
    t_clip, hw = wl["t"], wl["hw"]
    tower = vit.B200VisionTower(dict(CFG_7B), device=dev, return_dict=False)
    random_state_dict_gpu(tower, seed=0)

    # ---- this rank's share of one step: a list of (frames [T,H,W,3], per-frame shifts)
    layer = prompt_layer(hw)
    shifts_all = frame_shifts(t_clip)
    if wl.get("split"):                                   # one clip, contiguous temporal slices per rank
        t0s, t1s = vit.shard_slices(t_clip // 2, world, rank)
        clip_ids, f0, f1 = [0], 2 * t0s, 2 * t1s
    elif "batch" in wl:                                   # fixed batch sharded by clip
        clip_ids, f0, f1 = vit.shard_clips(wl["batch"], world, rank), 0, t_clip
    else:                                                 # one clip per rank
        clip_ids, f0, f1 = [rank], 0, t_clip
    t_local = f1 - f0
    frames_per_step_total = (wl["batch"] * t_clip if "batch" in wl else t_clip if wl.get("split") else t_clip * world)
    clips_host = [synthetic_frames(t_clip, hw, hw, clip_id=c)[f0:f1].contiguous().pin_memory() for c in clip_ids]
    clips_dev = [c.to(dev) for c in clips_host]
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=s, sy=s) for s in shifts_all[f0:f1]]
    overlay = vit.OverlaySpec.from_rgba(layer, ops, device=dev)
    grid = [[t_local // 2, hw // 14, hw // 14]]
    m = grid[0][0] * grid[0][1] * grid[0][2]
    n_local = len(clips_dev)
    rows_per_rank = [((vit.shard_slices(t_clip // 2, world, r)[1] - vit.shard_slices(t_clip // 2, world, r)[0]) * (hw // 28) ** 2
                      if wl.get("split") else m // 4) for r in range(world)]
    # two output sets (consecutive steps alternate) so a gather still in flight never aliases the next step's output
    outs = [[torch.empty(m // 4, CFG_7B["out_hidden_size"], dtype=torch.bfloat16, device=dev) for _ in range(max(n_local, 1))]
            for _ in range(2)]
    equal_rows = len(set(rows_per_rank)) == 1
    gather_lists = [[[torch.empty_like(outs[0][0]) for _ in range(world)] if (world > 1 and rank == 0 and equal_rows) else None
                     for _ in range(max(n_local, 1))] for _ in range(2)]
    do_gather = world > 1 and not os.environ.get("BENCH_NO_GATHER")

    pending = []
    step_no = [0]

    def gather_async(t, lst):
        if equal_rows:
            return dist.gather(t, lst, dst=0, async_op=True)   # merged tokens -> LLM rank (NCCL over NVLink)
        vit.gather_tokens(t, rows_per_rank, dst=0)
        return None

    def step_resident():
        """This rank's clips of one step through the path; with N > 1 the merged tokens of each clip go to rank 0 on
        NCCL's stream while the next clip is already being computed."""
        b = step_no[0] & 1
        step_no[0] += 1
        while len(pending) > n_local:                     # the set written two steps ago is about to be reused
            h = pending.pop(0)
            if h is not None:
                h.wait()                                  # stream-level wait, the host does not block
        for ci in range(n_local):
            tower.forward_frames(clips_dev[ci], overlay, out=outs[b][ci])
            if do_gather:
                pending.append(gather_async(outs[b][ci], gather_lists[b][ci]))

    def drain():
        while pending:
            h = pending.pop(0)
            if h is not None:
                h.wait()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- resident-input timing
    for _ in range(args.warmup - 1):
        step_resident()
    drain()
    torch.cuda.synchronize(dev)
    t_w = time.perf_counter()
    step_resident()
    drain()
    torch.cuda.synchronize(dev)
    est_total = (time.perf_counter() - t_w) * args.steps
    barrier()
    # Clocks.  N = 1: NVML (the recipe's nvidia-smi line) sampled in a separate process during the timed region.
    # N > 1: NVML reads while the ranks are coupled through NCCL inflate the step by 15-90 % (measured at N = 2, 30 steps:
    # 10.70 ms/step unsampled, 12.2-20.1 ms with 2-3 samples, whichever NVML call is made and whichever GPU is read), so
    # the SM clock of the timed region comes from an in-kernel probe instead (b200vit_clock_probe: SM cycles per
    # globaltimer nanosecond over 50 us, launched on a side stream every few steps), and throttle reasons / power are
    # read through NVML during an immediate untimed repeat of the same loop.
    use_probe = world > 1 and not os.environ.get("BENCH_NVML_IN_REGION")
    sampler = ClockSampler(int(os.environ.get("BENCH_SAMPLER_GPU", local_rank)),
                           interval=min(max(est_total / 5.0, 0.1), 2.0) if not use_probe else 0.1)
    n_probe = 6
    probe_every = max(args.steps // n_probe, 1)
    probe_buf = torch.zeros((args.steps // probe_every + 1, 2), dtype=torch.int64, device=dev)
    probe_stream = torch.cuda.Stream(dev, priority=-1)
    probes = [0]

    def maybe_probe(i):
        if use_probe and rank == 0 and i % probe_every == probe_every // 2:
            _lib.check(_lib.lib().b200vit_clock_probe(probe_buf[probes[0]].data_ptr(), 50000, probe_stream.cuda_stream), "probe")
            probes[0] += 1

    if rank == 0 and not use_probe and not os.environ.get("BENCH_NO_CLOCKS"):
        sampler.start()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident()
        maybe_probe(i)
    drain()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = None
    if not use_probe:
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    else:
        # untimed repeat of the same loop with the NVML sampler running: throttle reasons, power, NVML's own clock
        if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):
            sampler.start()
        barrier()
        r_wall0 = time.time()
        for _ in range(min(args.steps, 30)):
            step_resident()
        drain()
        barrier()
        r_wall1 = time.time()
        if rank == 0:
            nv = sampler.stop(r_wall0, r_wall1)
            pb = probe_buf[:probes[0]].cpu().numpy().astype(np.float64)
            mhz = [c / ns * 1e3 for c, ns in pb if ns > 0]
            clocks = {"sm_mhz": float(np.median(mhz)) if mhz else nv.get("sm_mhz"), "sm_max_mhz": nv.get("sm_max_mhz"),
                      "reasons": nv.get("reasons", []), "samples": len(mhz),
                      "source": "timed region: in-kernel probe (SM cycles / globaltimer ns over 50 us, every "
                                f"{probe_every} steps); reasons/power: NVML during an untimed repeat of the loop "
                                "(NVML reads inside the NCCL-coupled timed loop inflate the step time, see bench.py)",
                      "nvml_repeat": {k: nv.get(k) for k in ("sm_mhz", "power_w_max", "samples")}}

    # ---- host cost of enqueueing one clip: two forwards into an empty queue (no back-pressure from the GPU)
    t_h = time.perf_counter()
    for _ in range(2):
        tower.forward_frames(clips_dev[0], overlay, out=outs[0][0])
    host_ms = (time.perf_counter() - t_h) * 1e3 / 2
    torch.cuda.synchronize(dev)

    # ---- per-kernel breakdown (cudaEvent pairs around every launch; separate pass)
    tower.profile(grid, True)
    kinds = {}
    prof_reps = max(1, min(args.steps, 20) // max(n_local, 1))
    for _ in range(prof_reps):
        tower.forward_frames(clips_dev[0], overlay, out=outs[0][0])
        for k, (t_ms, n) in tower.profile_read(grid).items():
            a = kinds.setdefault(k, [0.0, 0])
            a[0] += t_ms
            a[1] += n
    tower.profile(grid, False)

    # ---- end-to-end through the public feed API (rga3_release_b200.ClipPipeline): pinned host uint8 frames in, host
    # embeddings out, H2D / D2H on their own streams, two clips in flight
    e2e_gather = (lambda t: dist.gather(t, gather_lists[0][0], dst=0)) if (do_gather and equal_rows) else \
                 ((lambda t: vit.gather_tokens(t, rows_per_rank, dst=0)) if do_gather else None)
    pipe = vit.ClipPipeline(tower, tuple(clips_host[0].shape), depth=2, to_host=True, after_forward=e2e_gather)

    def e2e_loop(n):
        for _ in range(n):
            for c in clips_host:
                pipe.submit(c, overlay)
        pipe.drain()                                   # the last results are in host memory

    e2e_loop(min(args.warmup, 3))
    barrier()
    f0e, f1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0e.record()
    e2e_loop(args.steps)
    f1e.record()
    barrier()
    ms2 = torch.tensor([f0e.elapsed_time(f1e)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms_total = float(ms2.item())

    if rank == 0:
        peaks = measured_peaks()
        flops_local, per_flops = algorithmic_flops(grid)        # one local clip (or slice range)
        total_flops_step = algorithmic_flops([[t_clip // 2, hw // 14, hw // 14]])[0] * (frames_per_step_total / t_clip)
        frames_total = frames_per_step_total * args.steps
        fps = frames_total / (ms_total * 1e-3)
        step_ms = ms_total / args.steps
        launches = tower.launches_per_forward(grid, with_frames=True) * n_local
        if kinds.get("qkv_rope_winattn", [0, 0])[1]:   # windowed layers: projection + window attention in one kernel
            n_win = CFG_7B["depth"] - len(CFG_7B["fullatt_block_indexes"])
            per_flops["qkv_rope_winattn"] = per_flops["qkv_rope"] * n_win / CFG_7B["depth"] + per_flops["attn_window"]
            per_flops["qkv_rope"] *= 1.0 - n_win / CFG_7B["depth"]
        breakdown = {k: round(v[0] / prof_reps * n_local, 4) for k, v in kinds.items() if v[1]}   # this GPU's clips of one step
        dom = max((k for k in per_flops if k in kinds and kinds[k][1]), key=lambda k: kinds[k][0])
        dom_ms = kinds[dom][0] / kinds[dom][1]                    # average launch duration
        dom_flops = per_flops[dom] / (kinds[dom][1] / prof_reps)  # algorithmic FLOPs per launch
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        traffic, traffic_src = measured_traffic(dom) if args.workload == "cfg2" else (None, None)
        tokens_per_frame = (hw // 28) ** 2 / 2.0
        line = {
            "metric": "vision_tower_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args.workload, world),
            "tokens_per_s": fps * tokens_per_frame,
            "tower_tflops": total_flops_step / (step_ms * 1e-3) / 1e12,
            "pct_bf16_peak_burst": total_flops_step / world / (step_ms * 1e-3) / 1e12 / peaks["bf16_burst"],
            "pct_bf16_peak_sustained": total_flops_step / world / (step_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
            "clocks": clocks,
            "e2e": {"value": frames_total / (e2e_ms_total * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(pipe.h2d_bytes) * n_local, "d2h_bytes_per_step": int(pipe.d2h_bytes) * n_local,
                    "ms_per_step": e2e_ms_total / args.steps},
            "gpu_launches": launches * args.steps, "host_enqueue_ms_per_step": host_ms * n_local,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
                         "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                         "flops_per_launch": dom_flops, "avg_launch_ms": dom_ms},
            "kernel_ms_per_step": breakdown,
        }
        if world == 1 and args.workload == "cfg2" and not args.no_gpu_baseline:
            # the same pixels the tower saw, in the HF processor's layout, for the stock GPU path
            pv = torch.empty(m, 1176, dtype=torch.bfloat16, device=dev)
            fr = _lib.Frames(clips_dev[0].data_ptr(), t_local, hw, hw)
            _lib.check(_lib.lib().b200vit_overlay_patchify(fr, overlay.to_c(t_local), 14, 2, 2, pv.data_ptr(),
                                                           torch.cuda.current_stream(dev).cuda_stream), "overlay_patchify")
            line["gpu_baseline"] = gpu_baseline(dev, pv, torch.tensor(grid, device=dev))
        if world == 1 and args.workload == "cfg2" and not args.no_cpu_baseline:
            stepf, nf, parts = reference_step_fn()
            stepf()
            t0 = time.perf_counter()
            stepf()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = cpu_baseline_entry(nf / dt, parts, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
