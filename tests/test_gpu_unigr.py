"""BASELINE config 5 boundary: the B200 tower inside the reference's own wrapper class.

/root/reference/model/qwen_2_5_vl_sam2.py::UniGRModel (:104) calls ``super().forward`` for prefill / generation
(:143-146, :182-200), i.e. HF Qwen2_5_VLForConditionalGeneration.forward -> get_video_features -> ``self.visual``.
/root/reference does not exist on the GPU box, so the reference class ran in the build container
(tests/golden/make_unigr_golden.py, CPU fp32) and its prefill logits are the fixture; here the same weights are rebuilt
from the same seed through the parent class, checked against the fixture, and then the tower is swapped."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import rga3_release_b200 as vit
from unigr_case import build_inputs, perturb_tower, unigr_config_kwargs

DEV = "cuda"


def test_prefill_logits_inside_the_unigr_wrapper(golden_dir):
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    z = np.load(os.path.join(golden_dir, "unigr_prefill.npz"))
    golden = torch.from_numpy(z["logits"])
    torch.manual_seed(0)
    model = Qwen2_5_VLForConditionalGeneration(Qwen2_5_VLConfig(**unigr_config_kwargs())).eval()
    perturb_tower(model)
    assert sum(p.numel() for p in model.parameters()) == int(z["n_params"])
    ids, types, pv, grid = build_inputs()
    kw = dict(input_ids=ids, pixel_values_videos=pv, video_grid_thw=grid, mm_token_type_ids=types, past_key_values=None)
    with torch.no_grad():
        stock_cpu = model(**kw).logits.float()
    # same class behaviour, same seed: the rebuilt model IS the reference wrapper's model (fp32 CPU, BLAS-order noise only)
    assert torch.allclose(stock_cpu, golden, rtol=1e-3, atol=1e-4), (stock_cpu - golden).abs().max()
    model = model.to(DEV)
    kw = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    tower = vit.install(model)
    assert isinstance(model.model.visual, vit.B200VisionTower) and tower.dtype == torch.float32
    with torch.no_grad():
        out = model(**kw).logits.float().cpu()
    a, b = out.double().flatten(), golden.double().flatten()
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    rel = ((a - b).abs().max() / b.abs().max()).item()
    assert cos >= 0.999 and rel <= 2e-2, (cos, rel)
    # the M-RoPE ids the model built internally are the ones mrope_position_ids returns for this transformers version
    pos, delta = vit.mrope_position_ids(kw["input_ids"], kw["mm_token_type_ids"], video_grid_thw=kw["video_grid_thw"],
                                        tokens_per_second=model.config.vision_config.tokens_per_second)
    want_pos, want_delta = model.model.get_rope_index(kw["input_ids"], mm_token_type_ids=kw["mm_token_type_ids"],
                                                      video_grid_thw=kw["video_grid_thw"])
    assert torch.equal(pos, want_pos) and torch.equal(delta, want_delta.to(delta.dtype))
    with torch.no_grad():
        out2 = model(position_ids=pos, **kw).logits.float().cpu()
    assert torch.equal(out2, out)
