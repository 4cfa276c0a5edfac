"""M-RoPE position ids for the LLM side of the splice: what HF ``get_rope_index`` computes for the
sequence the visual tokens are written into (reference call path: /root/reference/model/qwen_2_5_vl_sam2.py:182-200 ->
Qwen2_5_VLModel.forward -> get_rope_index, transformers modeling_qwen2_5_vl.py:1024-1133).

Two published behaviours exist and they DIFFER for every video grid with t > 1:
  * ``variant="4.49"`` -- the algorithm of the Qwen2.5-VL release that the reference pins (transformers 4.49.0.dev0,
    requirements.txt:25): inside a video, token (t, h, w) gets (t * interval, h, w) + offset with
    interval = second_per_grid_t * tokens_per_second truncated per token, the next text token continues at max + 1.
    That version is not installed here, so this branch is pinned only by its restatement in oracle/mrope_ref.py.
  * ``variant="5.5"`` -- transformers 5.5.0 as installed: one (start * interval) temporal index for the whole grid,
    heights ``arange(h).repeat_interleave(w * t)``, widths ``arange(w).repeat(h * t)``, the next segment continues at
    start + max(h, w) // merge.  Bit-exact against the installed ``get_rope_index`` (tests/test_mrope_cpu.py).
``variant="auto"`` picks by the installed transformers major version, so a tower dropped into a loaded model sees the
ids that model's own forward would have built.  Images (t = 1) are identical in both.
Host integer arithmetic (numpy), no kernel: L is a few thousand.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def _variant(variant: str) -> str:
    if variant != "auto":
        if variant not in ("4.49", "5.5"):
            raise ValueError("variant must be 'auto', '4.49' or '5.5'")
        return variant
    try:
        import transformers
        return "5.5" if int(transformers.__version__.split(".")[0]) >= 5 else "4.49"
    except Exception:
        return "4.49"


def _runs(types: np.ndarray):
    """(type, start, end) of every run of equal values (itertools.groupby in HF :1100-1104)."""
    if types.size == 0:
        return []
    cut = np.flatnonzero(np.diff(types)) + 1
    starts = np.concatenate([[0], cut])
    ends = np.concatenate([cut, [types.size]])
    return [(int(types[s]), int(s), int(e)) for s, e in zip(starts, ends)]


def mrope_position_ids(input_ids, mm_token_type_ids, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
                       attention_mask=None, spatial_merge_size: int = 2, tokens_per_second: int = 2,
                       variant: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
    """(position_ids [3, B, L] int64, rope_deltas [B, 1] int64) on ``input_ids``' device.
    ``mm_token_type_ids`` [B, L]: 0 text, 1 image, 2 video (HF 5.x processor output; for 4.49-style inputs build it
    from the placeholder ids: ``(ids == image_token_id) * 1 + (ids == video_token_id) * 2``)."""
    variant = _variant(variant)
    dev = input_ids.device if isinstance(input_ids, torch.Tensor) else torch.device("cpu")
    ids = np.asarray(torch.as_tensor(input_ids).cpu())
    types_all = np.asarray(torch.as_tensor(mm_token_type_ids).cpu()).astype(np.int64)
    am = None if attention_mask is None else np.asarray(torch.as_tensor(attention_mask).cpu()).astype(bool)
    grids = {1: iter(np.asarray(torch.as_tensor(image_grid_thw).cpu()).reshape(-1, 3).tolist()) if image_grid_thw is not None else None,
             2: iter(np.asarray(torch.as_tensor(video_grid_thw).cpu()).reshape(-1, 3).tolist()) if video_grid_thw is not None else None}
    if second_per_grid_ts is not None:
        spg = iter(np.asarray(torch.as_tensor(second_per_grid_ts).cpu()).reshape(-1).tolist())
    else:
        spg = None
    b, l = ids.shape
    # padding positions keep the initial fill: zeros in 5.5 (:1083), ones in 4.49
    pos = np.zeros((3, b, l), dtype=np.int64) if variant == "5.5" else np.ones((3, b, l), dtype=np.int64)
    deltas = []
    for bi in range(b):
        types = types_all[bi][am[bi]] if am is not None else types_all[bi]
        cur = 0
        chunks = []
        for kind, s, e in _runs(types):
            if kind == 0:
                chunks.append(np.arange(e - s, dtype=np.int64)[None, :].repeat(3, 0) + cur)
                cur += e - s
                continue
            t, h, w = next(grids[kind])
            gh, gw = h // spatial_merge_size, w // spatial_merge_size
            if variant == "5.5":
                # every vision run draws one entry, images included (HF :1119: next(second_per_grid_ts) per run)
                sec = next(spg) if spg is not None else 1
                interval = tokens_per_second * int(sec)
                n = gh * gw * t
                wpos = np.tile(np.arange(cur, cur + gw, dtype=np.int64), gh * t)
                hpos = np.repeat(np.arange(cur, cur + gh, dtype=np.int64), gw * t)
                tpos = np.full(n, cur, dtype=np.int64) * interval
                chunks.append(np.stack([tpos, hpos, wpos]))
                cur += max(h, w) // spatial_merge_size
            else:
                # published Qwen2.5-VL M-RoPE (transformers 4.49): seconds per grid only apply to videos
                sec = (next(spg) if spg is not None else 1.0) if kind == 2 else 0.0
                tt = (np.arange(t, dtype=np.float64)[:, None].repeat(gh * gw, 1) * float(sec) * tokens_per_second)
                tpos = tt.astype(np.int64).reshape(-1) if kind == 2 else np.zeros(t * gh * gw, dtype=np.int64)
                hpos = np.arange(gh, dtype=np.int64)[None, :, None].repeat(t, 0).repeat(gw, 2).reshape(-1)
                wpos = np.arange(gw, dtype=np.int64)[None, None, :].repeat(t, 0).repeat(gh, 1).reshape(-1)
                chunks.append(np.stack([tpos, hpos, wpos]) + cur)
                cur = int(chunks[-1].max()) + 1
        p = np.concatenate(chunks, axis=1) if chunks else np.zeros((3, 0), dtype=np.int64)
        if p.shape[1] != types.size:
            raise ValueError("vision runs do not match the grids: placeholder count differs from t * h * w / merge^2")
        if am is not None:
            pos[:, bi, am[bi]] = p
        else:
            pos[:, bi] = p
        # 5.5 subtracts the unpadded length (:1131), 4.49 the padded one
        deltas.append(int(p.max()) + 1 - (int(types.size) if variant == "5.5" else l) if p.size else 0)
    return (torch.from_numpy(pos).to(dev), torch.tensor(deltas, dtype=torch.int64, device=dev).unsqueeze(1))
