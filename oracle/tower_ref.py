"""fp32 restatement of the Qwen2.5-VL vision tower forward (plain torch, CPU).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Follows HF transformers 5.5.0 models/qwen2_5_vl/modeling_qwen2_5_vl.py
(``HF:modeling``), which the reference reaches via
/root/reference/model/qwen_2_5_vl_sam2.py:182-200 and :346-355:
  RMSNorm :57-71, MLP :77-88, PatchEmbed :91-114, rotary :117-130, :149-167,
  PatchMerger :133-146, attention :207-287, block :290-321, forward :455-518.
Works directly on an HF-style ``state_dict`` so it shares weights with the
real HF tower and with the CUDA module.  Pinned against the real HF tower in
tests/test_oracle_cpu.py and against tests/golden/tower_tiny.npz.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import index_ref


@dataclass
class TowerCfg:
    depth: int = 32
    hidden_size: int = 1280
    intermediate_size: int = 3420
    num_heads: int = 16
    out_hidden_size: int = 3584
    patch_size: int = 14
    temporal_patch_size: int = 2
    spatial_merge_size: int = 2
    window_size: int = 112
    in_channels: int = 3
    fullatt_block_indexes: List[int] = field(default_factory=lambda: [7, 15, 23, 31])

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    @classmethod
    def from_hf(cls, c):
        return cls(depth=c.depth, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                   num_heads=c.num_heads, out_hidden_size=c.out_hidden_size, patch_size=c.patch_size,
                   temporal_patch_size=c.temporal_patch_size, spatial_merge_size=c.spatial_merge_size,
                   window_size=c.window_size, in_channels=c.in_channels,
                   fullatt_block_indexes=list(c.fullatt_block_indexes))


def rmsnorm_ref(x, w, eps=1e-6):
    """HF:modeling:66-71."""
    x = x.float()
    var = x.pow(2).mean(-1, keepdim=True)
    return w.float() * (x * torch.rsqrt(var + eps))


def rope_ref(q, k, cos, sin):
    """HF:modeling:149-167 (rotate_half)."""
    def rot(x):
        h = x.shape[-1] // 2
        return torch.cat((-x[..., h:], x[..., :h]), dim=-1)
    cos, sin = cos.unsqueeze(-2), sin.unsqueeze(-2)
    return q * cos + rot(q) * sin, k * cos + rot(k) * sin


def varlen_attention_ref(q, k, v, cu_seqlens, scale):
    """HF:modeling:263-283 + :182-204: independent softmax(QK^T*scale)V per
    cu_seqlens segment; q,k,v [M, heads, head_dim] -> [M, heads*head_dim]."""
    m, nh, hd = q.shape
    out = torch.empty(m, nh, hd, dtype=torch.float32)
    cu = [int(c) for c in cu_seqlens]
    for s, e in zip(cu[:-1], cu[1:]):
        if e <= s:
            continue
        qs, ks, vs = (t[s:e].transpose(0, 1) for t in (q, k, v))
        att = torch.softmax(torch.matmul(qs, ks.transpose(1, 2)) * scale, dim=-1, dtype=torch.float32)
        out[s:e] = torch.matmul(att, vs).transpose(0, 1)
    return out.reshape(m, nh * hd)


def tower_forward_ref(sd: Dict[str, torch.Tensor], cfg: TowerCfg, pixel_values: torch.Tensor, grid_thw,
                      return_intermediates: bool = False):
    """HF:modeling:455-518.  Returns pooler_output [M/4, out_hidden] fp32 (and,
    optionally, a dict of intermediates keyed like the CUDA debug taps)."""
    grid = np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3)
    taps = {}
    x = pixel_values.float()
    m = x.shape[0]
    unit = cfg.spatial_merge_size ** 2
    w_pe = sd["patch_embed.proj.weight"].float().reshape(cfg.hidden_size, -1)
    x = x @ w_pe.t()                                                    # :106-114
    widx, _, cu_win = index_ref.window_index_ref(grid, cfg.window_size, cfg.spatial_merge_size, cfg.patch_size)
    cu_full = index_ref.cu_seqlens_ref(grid)
    rot = torch.from_numpy(index_ref.rope_table_ref(grid, cfg.head_dim, 10000.0, cfg.spatial_merge_size))
    widx_t = torch.from_numpy(widx)
    x = x.reshape(m // unit, unit, -1)[widx_t].reshape(m, -1)           # :478-481
    rot = rot.reshape(m // unit, unit, -1)[widx_t].reshape(m, -1)       # :482-484
    emb = torch.cat((rot, rot), dim=-1)
    cos, sin = emb.cos(), emb.sin()                                     # :485-486
    taps["patch_embed_reordered"] = x.clone()
    nh, hd = cfg.num_heads, cfg.head_dim
    for li in range(cfg.depth):
        p = f"blocks.{li}."
        cu = cu_full if li in cfg.fullatt_block_indexes else cu_win    # :498-502
        h = rmsnorm_ref(x, sd[p + "norm1.weight"])
        qkv = h @ sd[p + "attn.qkv.weight"].float().t() + sd[p + "attn.qkv.bias"].float()
        q, k, v = qkv.reshape(m, 3, nh, hd).permute(1, 0, 2, 3).unbind(0)   # :230-232
        q, k = rope_ref(q, k, cos, sin)
        a = varlen_attention_ref(q, k, v, cu, hd ** -0.5)
        a = a @ sd[p + "attn.proj.weight"].float().t() + sd[p + "attn.proj.bias"].float()
        x = x + a                                                       # :313
        h = rmsnorm_ref(x, sd[p + "norm2.weight"])
        g = h @ sd[p + "mlp.gate_proj.weight"].float().t() + sd[p + "mlp.gate_proj.bias"].float()
        u = h @ sd[p + "mlp.up_proj.weight"].float().t() + sd[p + "mlp.up_proj.bias"].float()
        d = (F.silu(g) * u) @ sd[p + "mlp.down_proj.weight"].float().t() + sd[p + "mlp.down_proj.bias"].float()
        x = x + d                                                       # :320
        if return_intermediates and li in (0, cfg.depth - 1):
            taps[f"block{li}"] = x.clone()
    taps["last_hidden_state"] = x
    h = rmsnorm_ref(x, sd["merger.ln_q.weight"]).reshape(m // unit, unit * cfg.hidden_size)   # :143-145
    h = F.gelu(h @ sd["merger.mlp.0.weight"].float().t() + sd["merger.mlp.0.bias"].float())
    h = h @ sd["merger.mlp.2.weight"].float().t() + sd["merger.mlp.2.bias"].float()
    rev = torch.from_numpy(index_ref.reverse_index_ref(widx))
    out = h[rev]                                                        # :512-513
    if return_intermediates:
        return out, taps
    return out
