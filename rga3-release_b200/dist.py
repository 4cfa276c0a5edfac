"""Multi-GPU plumbing for the visual path: shard independent clips (or temporal slices of one
clip) over ranks, run the tower locally, gather the merged tokens on the LLM rank.

The reference shards inference the same way -- one process per GPU, job i goes to rank
i % world (evaluation/videoinfer/run_inference_parallel.sh:21-29, inference_videoinfer.py:47-52)
-- and has no collective on the forward path; the gather below is the one exchange this
path adds.  torch.distributed is plumbing only (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_clips(n_clips: int, world: int, rank: int) -> List[int]:
    """Clip indices owned by `rank`: contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def shard_slices(grid_t: int, world: int, rank: int) -> Tuple[int, int]:
    """Temporal-slice range [t0, t1) of ONE clip owned by `rank` (slices are independent through
    all 32 layers: no attention segment spans slices, HF modeling :488-496)."""
    base, rem = divmod(grid_t, world)
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def gather_tokens(local: torch.Tensor, rows_per_rank: Sequence[int], dst: int = 0,
                  group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Gather ragged [rows_r, C] token blocks to `dst` in rank order.  Returns the concatenation on
    `dst`, None elsewhere.  Point-to-point (NCCL has no gatherv): dst posts one irecv per peer."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(rows_per_rank) == world and local.shape[0] == rows_per_rank[rank]
    if world == 1:
        return local
    if rank == dst:
        out = local.new_empty((sum(rows_per_rank), local.shape[1]))
        offs = [0]
        for r in rows_per_rank:
            offs.append(offs[-1] + r)
        reqs = []
        for r in range(world):
            view = out[offs[r]:offs[r + 1]]
            if r == dst:
                view.copy_(local)
            elif rows_per_rank[r]:
                reqs.append(dist.irecv(view, src=r, group=group))
        for q in reqs:
            q.wait()
        return out
    if local.shape[0]:
        dist.send(local.contiguous(), dst=dst, group=group)
    return None
