"""Multi-GPU correctness of the two sharding modes, run under torchrun on N >= 2 GPUs of one box:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py [--full]
  1. by clip (BASELINE cfg 3): a batch of clips sharded with shard_clips, merged tokens gathered to rank 0 ==
     rank 0 running every clip itself, BIT FOR BIT (same shapes -> same schedule -> same rounding).
  2. by temporal slice (one BASELINE cfg 4 clip): shard_slices, every rank runs frames [2 t0, 2 t1) with the matching
     prompt shifts, gathered tokens == the whole clip on one GPU within the path's tolerance (the slice count changes
     M, hence which tiles stream-K cuts in two).
--full uses the BASELINE shapes (8 x 32-frame 448^2 clips; one 64-frame 672^2 clip); default is a reduced clip count
and length at the same resolutions.  Prints MULTI_GPU_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import rga3_release_b200 as vit
from rga3_release_b200 import _lib


def main():
    full = "--full" in sys.argv
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tower = vit.B200VisionTower(dict(bench.CFG_7B), device=dev, return_dict=False)
    bench.random_state_dict_gpu(tower, seed=0)                 # same seed on every rank: replicated weights

    # ---- 1. by clip
    n_clips, t = (8, 32) if full else (max(world, 2) * 1, 8)
    layer = bench.prompt_layer(448)
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=s, sy=s) for s in bench.frame_shifts(t)]
    ov = vit.OverlaySpec.from_rgba(layer, ops, device=dev)
    mine = vit.shard_clips(n_clips, world, rank)
    outs = [tower.forward_frames(bench.synthetic_frames(t, 448, 448, c).to(dev), ov) for c in mine]
    rows = [len(vit.shard_clips(n_clips, world, r)) * outs[0].shape[0] for r in range(world)]
    local_cat = torch.cat(outs) if outs else torch.empty(0, 3584, dtype=torch.bfloat16, device=dev)
    gathered = vit.gather_tokens(local_cat, rows, dst=0)
    if rank == 0:
        want = torch.cat([tower.forward_frames(bench.synthetic_frames(t, 448, 448, c).to(dev), ov) for c in range(n_clips)])
        assert gathered.shape == want.shape and torch.equal(gathered, want), "clip-sharded gather differs from the single-GPU result"
        print(f"by clip: {n_clips} clips x {t} frames over {world} GPUs == single GPU, bit for bit", flush=True)

    # ---- 2. by temporal slice of one 672x672 clip
    t = 64 if full else 4 * world
    frames = bench.synthetic_frames(t, 672, 672, 7)
    layer = bench.prompt_layer(672)
    shifts = bench.frame_shifts(t)
    t0, t1 = vit.shard_slices(t // 2, world, rank)
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=s, sy=s) for s in shifts[2 * t0:2 * t1]]
    part = tower.forward_frames(frames[2 * t0:2 * t1].to(dev), vit.OverlaySpec.from_rgba(layer, ops, device=dev))
    rows = [(vit.shard_slices(t // 2, world, r)[1] - vit.shard_slices(t // 2, world, r)[0]) * 576 for r in range(world)]
    gathered = vit.gather_tokens(part, rows, dst=0)
    if rank == 0:
        ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=s, sy=s) for s in shifts]
        whole = tower.forward_frames(frames.to(dev), vit.OverlaySpec.from_rgba(layer, ops, device=dev))
        a, b = gathered.double().flatten(), whole.double().flatten()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        rel = ((a - b).abs().max() / b.abs().max()).item()
        assert gathered.shape == whole.shape and cos >= 0.9999 and rel <= 2e-2, (cos, rel)
        print(f"by slice: one {t}-frame 672x672 clip over {world} GPUs vs single GPU: cos {cos:.6f} rel {rel:.4f}", flush=True)
        print("MULTI_GPU_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
