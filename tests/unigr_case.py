"""The tiny UniGR-shaped case shared by tests/golden/make_unigr_golden.py (reference wrapper, CPU) and
tests/test_gpu_unigr.py (same weights rebuilt on the GPU box from the seed)."""
import torch

VID, IMG, VSTART, VEND = 990, 991, 992, 993


def unigr_config_kwargs():
    vision = dict(depth=2, hidden_size=160, intermediate_size=200, num_heads=2, out_hidden_size=128, window_size=112,
                  fullatt_block_indexes=[1])
    text = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, intermediate_size=256,
                vocab_size=1000, max_position_embeddings=4096, rope_scaling={"type": "mrope", "mrope_section": [4, 6, 6]})
    return dict(text_config=text, vision_config=vision, video_token_id=VID, image_token_id=IMG,
                vision_start_token_id=VSTART, vision_end_token_id=VEND)


def perturb_tower(model):
    """non-trivial biases / norm weights in the tower (HF initialises them to 0 / 1, which hides bugs)"""
    g = torch.Generator().manual_seed(2)
    visual = model.model.visual if hasattr(model.model, "visual") else model.visual
    with torch.no_grad():
        for n, p in visual.named_parameters():
            if n.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif "norm" in n or "ln_q" in n:
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))


def build_inputs():
    grid = torch.tensor([[2, 8, 12], [1, 6, 10]])
    ids, types = [1, 2], [0, 0]
    for t, h, w in grid.tolist():
        n = t * h * w // 4
        ids += [VSTART] + [VID] * n + [VEND]
        types += [0] + [2] * n + [0]
    ids += [5, 6, 7]
    types += [0, 0, 0]
    m = int(grid.prod(-1).sum())
    pv = torch.randn(m, 1176, generator=torch.Generator().manual_seed(1))
    return torch.tensor([ids]), torch.tensor([types], dtype=torch.int32), pv, grid
