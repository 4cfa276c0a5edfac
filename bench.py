#!/usr/bin/env python
"""Benchmark of the RGA3 visual path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one synthetic clip of BASELINE config 2
(16 frames, 448x448, box + mask prompt overlay on every frame, per-frame shift):
overlay -> normalise -> patchify -> 32-layer Qwen2.5-VL-7B-shaped vision tower ->
merged visual embeddings [2048, 3584].  N > 1 (torchrun, one rank per GPU): every
rank processes its own clip per step (weak scaling) and the merged tokens are
gathered to rank 0 with NCCL inside the timed region.

`value`  : frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through the public module call with HOST (pinned) uint8 frames in
           and HOST embeddings out, H2D/D2H inside the timed region.
`--impl reference`: the reference's own CPU implementation of the path (PIL overlay ->
           HF Qwen2VLVideoProcessor -> HF tower fp32 eager) on this box's host cores, on a
           bounded sample of the same clip.
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

T_FRAMES, H, W = 16, 448, 448
CFG_7B = dict(depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=3584,
              window_size=112, fullatt_block_indexes=[7, 15, 23, 31])
WORKLOAD = "cfg2: one 16-frame 448x448 clip, STOM box+mask overlay on all 16 frames, grid_thw=[8,32,32], 8192 patches -> 2048 merged tokens"


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


# ----------------------------------------------------------------------------- synthetic workload
def synthetic_frames(t, h, w, clip_id=0):
    g = torch.Generator().manual_seed(1000 + clip_id)
    noise = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, generator=g)
    yy = torch.linspace(0, 1, h).view(1, h, 1, 1)
    xx = torch.linspace(0, 1, w).view(1, 1, w, 1)
    tt = torch.linspace(0, 1, max(t, 2))[:t].view(t, 1, 1, 1)
    smooth = (127.5 + 127.5 * torch.sin(6.28318 * (yy * 1.5 + xx * 0.75 + tt))).expand(t, h, w, 3)
    return ((noise.float() + smooth) * 0.5).round().clamp(0, 255).to(torch.uint8)


def prompt_layer():
    """cfg 2 prompt: red box outline (112,96,335,351) width 4 alpha 200 + lime disc r=80 alpha 100."""
    from PIL import Image, ImageDraw
    vip = Image.new("RGBA", (W, H), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    d.ellipse([(224 - 80, 224 - 80), (224 + 80, 224 + 80)], fill=(0, 255, 0, 100))
    d.rectangle([(112, 96), (335, 351)], outline=(255, 0, 0, 200), width=4)
    return np.array(vip)


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
    tower._invalidate()


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
def reference_step_fn(sample_frames=2):
    """The reference's CPU path on a bounded sample (first `sample_frames` frames = whole temporal slices, the
    tower's independent unit): PIL alpha_composite overlay -> HF Qwen2VLVideoProcessor -> HF tower fp32 eager."""
    from PIL import Image
    from oracle import hf_ref
    torch.set_num_threads(os.cpu_count() or 1)
    model, cfg, _ = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.float32, attn="eager")
    frames = synthetic_frames(T_FRAMES, H, W, 0)[:sample_frames].numpy()
    layer = prompt_layer()

    def step():
        comp = []
        for i in range(sample_frames):
            shifted = np.roll(layer, (i - 8, i - 8), axis=(0, 1))  # integer translate (layer is clear near the border)
            pil = Image.alpha_composite(Image.fromarray(frames[i], "RGB").convert("RGBA"), Image.fromarray(shifted, "RGBA"))
            comp.append(np.array(pil.convert("RGB")))
        pv, grid = hf_ref.hf_patchify(np.stack(comp))
        return hf_ref.hf_forward(model, pv, grid)
    return step, sample_frames


def run_reference(args, rank, world):
    if rank != 0:
        return
    step, nf = reference_step_fn(2)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = nf * args.steps / dt
    line = {"impl": "reference", "metric": "vision_tower_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_sample": f"{nf} of 16 frames per step (one temporal slice, grid_thw=[1,32,32])"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "reference",
                             "sample": f"{nf}-frame slice of the cfg2 clip per step: PIL overlay + HF video processor + HF tower fp32 eager on CPU"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tokens_per_s": fps * 128.0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("BENCH_STREAMS", "1")),
                    help="clips in flight per GPU (CUDA streams, one workspace each); measured: 2 is 2.7 % slower than 1")
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

    tower = vit.B200VisionTower(dict(CFG_7B), device=dev, return_dict=False)
    random_state_dict_gpu(tower, seed=0)
    grid = [[T_FRAMES // 2, H // 14, W // 14]]
    m = grid[0][0] * grid[0][1] * grid[0][2]
    layer = prompt_layer()
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i - 8, sy=i - 8) for i in range(T_FRAMES)]
    overlay = vit.OverlaySpec.from_rgba(layer, ops, device=dev)
    frames_host = synthetic_frames(T_FRAMES, H, W, clip_id=rank).pin_memory()
    frames_dev = frames_host.to(dev)
    out = torch.empty(m // 4, CFG_7B["out_hidden_size"], dtype=torch.bfloat16, device=dev)
    outs = [out, torch.empty_like(out)]
    gather_list = [torch.empty_like(out) for _ in range(world)] if (world > 1 and rank == 0) else None
    gather_lists = [gather_list, [torch.empty_like(out) for _ in range(world)] if gather_list is not None else None]

    do_gather = world > 1 and not os.environ.get("BENCH_NO_GATHER")
    gmode = os.environ.get("BENCH_GATHER", "gather")
    ag_buf = torch.empty(world * out.shape[0], out.shape[1], dtype=out.dtype, device=dev) if world > 1 else None

    def gather_out(t):
        if gmode == "gather":
            dist.gather(t, gather_list, dst=0)  # merged tokens -> LLM rank (NCCL over NVLink)
        elif gmode == "p2p":
            vit.gather_tokens(t, [t.shape[0]] * world, dst=0)
        else:
            dist.all_gather_into_tensor(ag_buf, t)

    pending = [None, None]
    step_no = [0]
    n_streams = max(1, min(args.streams, 2))
    main_stream = torch.cuda.current_stream(dev)
    side = [torch.cuda.Stream(dev) for _ in range(n_streams)] if n_streams > 1 else [main_stream]

    def step_resident():
        """One clip through the path.  Consecutive clips alternate between `n_streams` CUDA streams (one workspace
        and one output buffer each), so one clip's kernel tails / set-up overlap the other clip's kernels; with
        N > 1 the merged tokens go to rank 0 on NCCL's stream while the next clip is already being computed."""
        b = step_no[0] & 1
        step_no[0] += 1
        st = side[b % n_streams]
        with torch.cuda.stream(st):
            if pending[b] is not None:
                pending[b].wait()          # stream-level wait, the host does not block
                pending[b] = None
            tower.forward_frames(frames_dev, overlay, out=outs[b], slot=b % n_streams)
            if do_gather:
                if gmode == "gather":
                    pending[b] = dist.gather(outs[b], gather_lists[b], dst=0, async_op=True)
                else:
                    gather_out(outs[b])

    def fork():
        for st in side:
            if st is not main_stream:
                st.wait_stream(main_stream)

    def drain():
        for b in (0, 1):
            st = side[b % n_streams]
            with torch.cuda.stream(st):
                if pending[b] is not None:
                    pending[b].wait()
                    pending[b] = None
        for st in side:
            if st is not main_stream:
                main_stream.wait_stream(st)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- resident-input timing
    fork()
    for _ in range(args.warmup - 1):
        step_resident()
    drain()
    torch.cuda.synchronize(dev)
    t_w = time.perf_counter()
    fork()
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
    fork()
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
        fork()
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

    # ---- host cost of enqueueing one step: two steps into an empty queue (no back-pressure from the GPU)
    t_h = time.perf_counter()
    for _ in range(2):
        tower.forward_frames(frames_dev, overlay, out=out)
    host_ms = (time.perf_counter() - t_h) * 1e3 / 2
    torch.cuda.synchronize(dev)

    # ---- per-kernel breakdown (cudaEvent pairs around every launch; separate pass over the same K steps)
    tower.profile(grid, True)
    kinds = {}
    for _ in range(args.steps):
        tower.forward_frames(frames_dev, overlay, out=out)
        for k, (t_ms, n) in tower.profile_read(grid).items():
            a = kinds.setdefault(k, [0.0, 0])
            a[0] += t_ms
            a[1] += n
    tower.profile(grid, False)

    # ---- end-to-end through the public feed API (rga3_release_b200.ClipPipeline): pinned host uint8 frames in, host
    # embeddings out, H2D / D2H on their own streams, two clips in flight
    pipe = vit.ClipPipeline(tower, tuple(frames_host.shape), depth=2, to_host=True,
                            after_forward=(gather_out if do_gather else None))

    def e2e_loop(n):
        for _ in range(n):
            pipe.submit(frames_host, overlay)
        pipe.drain()                                   # the last results are in host memory

    e2e_loop(args.warmup)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(args.steps)
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms_total = float(ms2.item())

    if rank == 0:
        peaks = measured_peaks()
        total_flops, per_flops = algorithmic_flops(grid)
        frames_total = T_FRAMES * args.steps * world
        fps = frames_total / (ms_total * 1e-3)
        step_ms = ms_total / args.steps
        launches = tower.launches_per_forward(grid, with_frames=True)
        breakdown = {k: round(v[0] / args.steps, 4) for k, v in kinds.items() if v[1]}
        dom = max((k for k in per_flops if k in kinds and kinds[k][1]), key=lambda k: kinds[k][0])
        dom_ms = kinds[dom][0] / kinds[dom][1]                    # average launch duration
        dom_flops = per_flops[dom] / (kinds[dom][1] / args.steps)  # algorithmic FLOPs per launch
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        line = {
            "metric": "vision_tower_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "weights": "random-init Qwen2.5-VL-7B vision tower shape (676.6M params)",
                       "parallelism": f"clip-sharded dp{world}, merged tokens gathered to rank 0" if world > 1 else "single GPU",
                       "clips_in_flight_per_gpu": n_streams,
                       "l2": "per-step working set (1.35 GB bf16 weights + 0.3 GB activations) exceeds the 126 MB L2; no flush needed"},
            "tokens_per_s": fps * (m // 4) / T_FRAMES,
            "tower_tflops": total_flops * world / (step_ms * 1e-3) / 1e12,
            "pct_bf16_peak_burst": total_flops / (step_ms * 1e-3) / 1e12 / peaks["bf16_burst"],
            "pct_bf16_peak_sustained": total_flops / (step_ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
            "clocks": clocks,
            "e2e": {"value": frames_total / (e2e_ms_total * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
                    "ms_per_step": e2e_ms_total / args.steps},
            "gpu_launches": launches * args.steps, "host_enqueue_ms_per_step": host_ms,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the ncu --set full
                         # capture in profiles/r01_ncu_gateup_gemm_full.csv (algorithmic minimum 95 MB: A 21 + W 17.7 +
                         # out 56.6; part of the output is still in L2 when the kernel ends)
                         "traffic": 57.4e6 if dom == "gateup_swiglu" else None, "traffic_unit": "bytes/launch",
                         "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                         "flops_per_launch": dom_flops, "avg_launch_ms": dom_ms},
            "kernel_ms_per_step": breakdown,
        }
        if world == 1 and not args.no_cpu_baseline:
            stepf, nf = reference_step_fn(2)
            stepf()
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                stepf()
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": nf / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "reference",
                                    "sample": f"{nf}-frame slice (grid_thw=[1,32,32]) of the cfg2 clip, {reps} timed calls after 1 warm-up: "
                                              "PIL overlay + HF video processor + HF tower fp32 eager on the host cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
