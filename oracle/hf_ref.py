"""Builders for the REAL reference implementation of the path (third-party
HuggingFace code the reference repo imports) with the seeded weights used by
the tests and the bench.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference's tower is ``Qwen2_5_VisionTransformerPretrainedModel``
(transformers; pinned 4.49.0.dev0 at /root/reference/requirements.txt:25,
5.5.0 in this image), instantiated by ``UniGRModel.from_pretrained`` at
/root/reference/app.py:50-56 and called at
/root/reference/model/qwen_2_5_vl_sam2.py:182-200.  ``transformers`` is part of
the image (site-packages), so this module also works on the GPU box; it never
reads /root/reference.
"""
from __future__ import annotations

from typing import Dict

import torch

from .tower_ref import TowerCfg

CFG_7B = dict(depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=3584,
              window_size=112, fullatt_block_indexes=[7, 15, 23, 31])
# small shapes that keep head_dim = 80 (the kernels' specialisation) for fast CPU checks
CFG_TINY = dict(depth=2, hidden_size=160, intermediate_size=212, num_heads=2, out_hidden_size=96,
                window_size=112, fullatt_block_indexes=[1])
CFG_SMALL = dict(depth=4, hidden_size=320, intermediate_size=428, num_heads=4, out_hidden_size=256,
                 window_size=112, fullatt_block_indexes=[1, 3])


def state_dict_shapes(cfg: TowerCfg) -> Dict[str, tuple]:
    """state_dict layout of the HF tower (probed from HF)."""
    d, i, o = cfg.hidden_size, cfg.intermediate_size, cfg.out_hidden_size
    u = cfg.spatial_merge_size ** 2
    shapes = {"patch_embed.proj.weight": (d, cfg.in_channels, cfg.temporal_patch_size, cfg.patch_size, cfg.patch_size)}
    for li in range(cfg.depth):
        p = f"blocks.{li}."
        shapes.update({
            p + "norm1.weight": (d,), p + "norm2.weight": (d,),
            p + "attn.qkv.weight": (3 * d, d), p + "attn.qkv.bias": (3 * d,),
            p + "attn.proj.weight": (d, d), p + "attn.proj.bias": (d,),
            p + "mlp.gate_proj.weight": (i, d), p + "mlp.gate_proj.bias": (i,),
            p + "mlp.up_proj.weight": (i, d), p + "mlp.up_proj.bias": (i,),
            p + "mlp.down_proj.weight": (d, i), p + "mlp.down_proj.bias": (d,),
        })
    shapes.update({
        "merger.ln_q.weight": (d,),
        "merger.mlp.0.weight": (u * d, u * d), "merger.mlp.0.bias": (u * d,),
        "merger.mlp.2.weight": (o, u * d), "merger.mlp.2.bias": (o,),
    })
    return shapes


def make_state_dict(cfg: TowerCfg, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic random-init weights (CPU generator, so identical on every
    machine): linear/conv weights ~ N(0, 0.02) like HF's initializer_range,
    biases ~ N(0, 0.02) and norm weights ~ 1 + 0.1 N(0,1) so that zero biases /
    unit norms cannot hide bugs."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_shapes(cfg).items():
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("ln_q.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t.to(dtype)
    return sd


def build_hf_tower(cfg_kwargs=None, seed: int = 0, dtype=torch.float32, attn="eager", device="cpu"):
    """The real HF tower with ``make_state_dict`` weights loaded."""
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel
    kw = dict(CFG_7B if cfg_kwargs is None else cfg_kwargs)
    hf_cfg = Qwen2_5_VLVisionConfig(**kw)
    hf_cfg._attn_implementation = attn
    with torch.device("meta"):
        model = Qwen2_5_VisionTransformerPretrainedModel._from_config(hf_cfg)
    model = model.to_empty(device="cpu")
    cfg = TowerCfg.from_hf(hf_cfg)
    sd = make_state_dict(cfg, seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("inv_freq" in k for k in missing), missing
    # inv_freq is a non-persistent buffer: recompute (HF:modeling:117-130)
    hd = cfg.head_dim // 2
    model.rotary_pos_emb.inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float) / hd))
    model = model.to(dtype=dtype, device=device).eval()
    model.rotary_pos_emb.inv_freq = model.rotary_pos_emb.inv_freq.float()
    return model, cfg, sd


@torch.no_grad()
def hf_forward(model, pixel_values, grid_thw):
    out = model(pixel_values, grid_thw=grid_thw)
    return out.pooler_output if hasattr(out, "pooler_output") else out


def hf_video_processor():
    from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor
    return Qwen2VLVideoProcessor()


def hf_patchify(frames_u8_thwc):
    """frames [T,H,W,3] uint8 numpy/torch -> (pixel_values fp32 [M,1176], grid_thw)
    through the real HF video processor (no resize)."""
    proc = hf_video_processor()
    fr = torch.as_tensor(frames_u8_thwc).permute(0, 3, 1, 2).contiguous()
    out = proc(videos=[fr], do_resize=False, return_tensors="pt", do_sample_frames=False)
    return out["pixel_values_videos"], out["video_grid_thw"]


def synthetic_frames(t, h, w, clip_id=0):
    """SURVEY.md 8d synthetic clip: seeded uint8 noise blended with a smooth
    component so values are not pure white noise."""
    g = torch.Generator().manual_seed(1000 + clip_id)
    noise = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, generator=g)
    yy = torch.linspace(0, 1, h).view(1, h, 1, 1)
    xx = torch.linspace(0, 1, w).view(1, 1, w, 1)
    tt = torch.linspace(0, 1, max(t, 2))[:t].view(t, 1, 1, 1)
    smooth = (127.5 + 127.5 * torch.sin(6.28318 * (yy * 1.5 + xx * 0.75 + tt))).expand(t, h, w, 3)
    return ((noise.float() + smooth) * 0.5).round().clamp(0, 255).to(torch.uint8)
