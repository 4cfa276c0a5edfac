"""Prefill logits of the REFERENCE's own wrapper class: /root/reference/model/qwen_2_5_vl_sam2.py::UniGRModel is
imported unmodified (qwen_vl_utils stubbed, SURVEY.md 8c recipe; train_mask_decoder=True skips the SAM2 modules,
:111-115), built from a tiny seeded config on the CPU in fp32 and called the way generation calls it
(``model(input_ids=..., pixel_values_videos=..., video_grid_thw=..., mm_token_type_ids=..., past_key_values=None)`` ->
``super().forward``, :143-146).  The GPU test rebuilds the same weights from the same seed WITHOUT /root/reference
(absent on the GPU box), checks the stock HF logits against this file, then swaps in the B200 tower.
  python tests/golden/make_unigr_golden.py   -> tests/golden/unigr_prefill.npz"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from unigr_case import build_inputs, perturb_tower, unigr_config_kwargs  # noqa: E402


def main():
    sys.modules["qwen_vl_utils"] = types.SimpleNamespace(process_vision_info=None)
    sys.path.insert(0, "/root/reference")
    from model.qwen_2_5_vl_sam2 import UniGRConfig, UniGRModel
    torch.manual_seed(0)
    cfg = UniGRConfig(train_mask_decoder=True, **unigr_config_kwargs())
    model = UniGRModel(cfg).eval()
    perturb_tower(model)
    ids, types_, pv, grid = build_inputs()
    with torch.no_grad():
        out = model(input_ids=ids, pixel_values_videos=pv, video_grid_thw=grid, mm_token_type_ids=types_, past_key_values=None)
    logits = out.logits.float().numpy()
    n_params = sum(p.numel() for p in model.parameters())
    np.savez_compressed(os.path.join(HERE, "unigr_prefill.npz"), logits=logits, n_params=np.array(n_params),
                        cls=np.array(f"{type(model).__module__}.{type(model).__name__}"))
    print("logits", logits.shape, "params", n_params, "class", type(model).__mro__[:2])


if __name__ == "__main__":
    main()
