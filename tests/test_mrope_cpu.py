"""M-RoPE position ids (rga3-release_b200/mrope.py): the 5.5 rule bit-exact against the installed transformers'
get_rope_index, the 4.49 rule (the reference's pinned version, not installable here) against its restatement."""
import numpy as np
import pytest
import torch

import rga3_release_b200 as vit
from oracle import mrope_ref

VID, IMG = 990, 991


def make_batch(rng, n_seq, with_images=True):
    seqs, types, vgrids, igrids = [], [], [], []
    for _ in range(n_seq):
        ids, ty = [], []
        for _ in range(int(rng.integers(1, 4))):
            n_txt = int(rng.integers(0, 6))
            ids += rng.integers(1, 900, n_txt).tolist()
            ty += [0] * n_txt
            if with_images and rng.random() < 0.4:
                g = [1, int(rng.integers(1, 5)) * 2, int(rng.integers(1, 5)) * 2]
                igrids.append(g)
                n = g[1] * g[2] // 4
                ids += [IMG] * n
                ty += [1] * n
            else:
                g = [int(rng.integers(1, 5)), int(rng.integers(1, 5)) * 2, int(rng.integers(1, 5)) * 2]
                vgrids.append(g)
                n = g[0] * g[1] * g[2] // 4
                ids += [VID] * n
                ty += [2] * n
            ids += [5]
            ty += [0]
        seqs.append(ids)
        types.append(ty)
    L = max(len(s) for s in seqs)
    ids = torch.zeros(n_seq, L, dtype=torch.long)
    ty = torch.zeros(n_seq, L, dtype=torch.int32)
    am = torch.zeros(n_seq, L, dtype=torch.long)
    for i, (s, t) in enumerate(zip(seqs, types)):      # left padding, as generation uses
        ids[i, L - len(s):] = torch.tensor(s)
        ty[i, L - len(s):] = torch.tensor(t, dtype=torch.int32)
        am[i, L - len(s):] = 1
    return ids, ty, am, igrids, vgrids, types


def hf_model():
    from transformers import Qwen2_5_VLConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VLModel
    cfg = Qwen2_5_VLConfig(
        text_config=dict(hidden_size=64, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1,
                         intermediate_size=64, vocab_size=1000, rope_scaling={"type": "mrope", "mrope_section": [4, 6, 6]}),
        vision_config=dict(depth=1, hidden_size=160, intermediate_size=64, num_heads=2, out_hidden_size=64,
                           fullatt_block_indexes=[0]),
        video_token_id=VID, image_token_id=IMG)
    return Qwen2_5_VLModel(cfg)


def test_variant_55_matches_installed_transformers():
    import transformers
    if int(transformers.__version__.split(".")[0]) < 5:
        pytest.skip("installed transformers is not 5.x")
    model = hf_model()
    rng = np.random.default_rng(0)
    for trial in range(25):
        ids, ty, am, ig, vg, _ = make_batch(rng, int(rng.integers(1, 4)))
        n_runs = len(ig) + len(vg)
        spg = torch.tensor(rng.integers(1, 4, n_runs).tolist(), dtype=torch.float32) if trial % 2 else None
        kw = dict(image_grid_thw=torch.tensor(ig) if ig else None, video_grid_thw=torch.tensor(vg) if vg else None,
                  second_per_grid_ts=spg, attention_mask=am)
        want_pos, want_delta = model.get_rope_index(ids, mm_token_type_ids=ty, **kw)
        got_pos, got_delta = vit.mrope_position_ids(ids, ty, spatial_merge_size=2,
                                                    tokens_per_second=model.config.vision_config.tokens_per_second,
                                                    variant="5.5", **kw)
        assert torch.equal(got_pos, want_pos) and torch.equal(got_delta, want_delta.to(got_delta.dtype))
        assert torch.equal(vit.mrope_position_ids(ids, ty, variant="auto", tokens_per_second=model.config.vision_config.tokens_per_second,
                                                  **kw)[0], want_pos)


@pytest.mark.parametrize("variant", ["4.49", "5.5"])
def test_variants_match_their_restatements(variant):
    rng = np.random.default_rng(1)
    ref = mrope_ref.rope_index_449_ref if variant == "4.49" else mrope_ref.rope_index_55_ref
    for trial in range(25):
        ids, ty, am, ig, vg, types = make_batch(rng, 1, with_images=variant == "4.49" or trial % 2 == 0)
        spg_v = rng.uniform(0.2, 2.5, len(vg)).tolist() if variant == "4.49" else rng.integers(1, 4, len(ig) + len(vg)).tolist()
        got, delta = vit.mrope_position_ids(ids, ty, image_grid_thw=ig or None, video_grid_thw=vg or None,
                                            second_per_grid_ts=spg_v, variant=variant)
        want = ref(types[0], ig, vg, spg_v)
        assert got[:, 0].tolist() == want
        assert int(delta[0, 0]) == max(max(w) for w in want) + 1 - len(types[0])


def test_known_answer_from_the_hf_docstring():
    """HF's worked example (modeling :1036-1049): fps 1, tokens_per_second 25, temporal patch 2 -> seconds per grid 2:
    3 temporal x 2 x 2 patches after merging: temporal ids 0, 50, 100 under the published (4.49) rule."""
    ty = torch.tensor([[2] * 12 + [0] * 3])
    pos, delta = vit.mrope_position_ids(torch.zeros(1, 15, dtype=torch.long), ty, video_grid_thw=[[3, 4, 4]],
                                        second_per_grid_ts=[2.0], tokens_per_second=25, variant="4.49")
    assert pos[0, 0].tolist() == [0] * 4 + [50] * 4 + [100] * 4 + [101, 102, 103]
    assert pos[1, 0].tolist() == [0, 0, 1, 1] * 3 + [101, 102, 103]
    assert pos[2, 0].tolist() == [0, 1, 0, 1] * 3 + [101, 102, 103]
    assert int(delta) == 104 - 15


def test_grid_mismatch_is_loud():
    with pytest.raises(ValueError):
        vit.mrope_position_ids(torch.zeros(1, 6, dtype=torch.long), torch.tensor([[2] * 5 + [0]]), video_grid_thw=[[1, 4, 4]])
