"""2-GPU diagnostic: where does the time go when a gather follows every forward?"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import rga3_release_b200 as vit
from rga3_release_b200 import _lib

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
tower = vit.B200VisionTower(dict(bench.CFG_7B), device=dev, return_dict=False)
bench.random_state_dict_gpu(tower)
frames = bench.synthetic_frames(16, 448, 448, rank).to(dev)
ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i - 8, sy=i - 8) for i in range(16)]
ov = vit.OverlaySpec.from_rgba(bench.prompt_layer(), ops, device=dev)
out = torch.empty(2048, 3584, dtype=torch.bfloat16, device=dev)
gl = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
mode = sys.argv[1] if len(sys.argv) > 1 else "gather"
for _ in range(4):
    tower.forward_frames(frames, ov, out=out); dist.gather(out, gl, dst=0)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
K = 20
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
host = []
t0 = time.perf_counter()
for i in range(K):
    ev[i][0].record()
    h0 = time.perf_counter()
    tower.forward_frames(frames, ov, out=out)
    h1 = time.perf_counter()
    ev[i][1].record()
    if mode == "gather":
        dist.gather(out, gl, dst=0)
    h2 = time.perf_counter()
    ev[i][2].record()
    host.append((h1 - h0, h2 - h1))
torch.cuda.synchronize()
t1 = time.perf_counter()
fwd = sum(ev[i][0].elapsed_time(ev[i][1]) for i in range(K)) / K
gat = sum(ev[i][1].elapsed_time(ev[i][2]) for i in range(K)) / K
tot = ev[0][0].elapsed_time(ev[K - 1][2]) / K
print(f"rank {rank} mode {mode}: gpu fwd {fwd:.2f} ms, gather {gat:.2f} ms, per-step {tot:.2f} ms | host fwd {1e3*sum(h[0] for h in host)/K:.2f} ms, host gather {1e3*sum(h[1] for h in host)/K:.2f} ms, wall {1e3*(t1-t0)/K:.2f}", flush=True)
dist.destroy_process_group()
