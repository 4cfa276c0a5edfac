"""Whole-path parity on the B200: the CUDA tower (through B200VisionTower ->
b200vit_forward) against the fp32 oracle and the real HF tower on identical
weights and inputs.  Tolerance from BASELINE.json north_star: cosine >= 0.999
and max|d| / max|ref| <= 2e-2 (bf16 against the fp32 reference)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import rga3_release_b200 as vit
from rga3_release_b200 import _lib
from oracle import hf_ref, index_ref, overlay_ref, patchify_ref, tower_ref

DEV = "cuda"
COS_MIN, REL_MAX = 0.999, 2e-2


def parity(out, ref):
    out, ref = out.double().cpu().flatten(), ref.double().cpu().flatten()   # fp64: fp32 sums over 7M elements drift
    cos = torch.nn.functional.cosine_similarity(out, ref, dim=0).item()
    rel = ((out - ref).abs().max() / ref.abs().max()).item()
    return cos, rel


def make_tower(cfg_kwargs, seed=0, **kw):
    cfg = tower_ref.TowerCfg(**cfg_kwargs)
    sd = hf_ref.make_state_dict(cfg, seed)
    t = vit.B200VisionTower(dict(cfg_kwargs), device=DEV, return_dict=False, **kw)
    t.load_state_dict(sd)
    return t, cfg, sd


@pytest.mark.parametrize("grid", [[[2, 8, 12]], [[1, 6, 10], [2, 4, 8]], [[1, 2, 2]], [[3, 18, 14]],
                                  [[2, 8, 8]], [[1, 16, 8], [2, 8, 16]]])   # the last two: every window 64 rows -> fused attention path
def test_tower_tiny_vs_oracle(grid):
    t, cfg, sd = make_tower(hf_ref.CFG_TINY, output_fp32=True)
    m = sum(a * b * c for a, b, c in grid)
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(5))
    ref, taps = tower_ref.tower_forward_ref(sd, cfg, x, grid, return_intermediates=True)
    out = t(x.to(DEV), torch.tensor(grid))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)
    out2 = t(x.to(DEV).to(torch.bfloat16), torch.tensor(grid), output_last_hidden_state=True)
    # return_dict=False -> tensor; ask again through the 5.x style object
    t.return_dict = True
    o = t(x.to(DEV), torch.tensor(grid), output_last_hidden_state=True)
    cos, rel = parity(o.last_hidden_state, taps["last_hidden_state"])
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_tower_golden_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "tower_tiny.npz"))
    t, cfg, sd = make_tower(hf_ref.CFG_TINY, output_fp32=True)
    grid = z["grid_thw"]
    m = int(np.prod(grid, axis=1).sum())
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(int(z["x_seed"])))
    out = t(x.to(DEV), torch.from_numpy(grid))
    cos, rel = parity(out, torch.from_numpy(z["pooler_output"]))
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_plan_indices_bitexact():
    t = vit.B200VisionTower(dict(hf_ref.CFG_TINY), device=DEV)
    for grid in ([[2, 8, 12]], [[8, 32, 32]], [[1, 6, 10], [2, 18, 14]], [[2, 48, 48]]):
        p = t.plan_for(grid)
        wi, raw, cu = index_ref.window_index_ref(grid)
        assert np.array_equal(p.get(_lib.PLAN_WINDOW_INDEX, np.int64), wi)
        assert np.array_equal(p.get(_lib.PLAN_REVERSE_INDEX, np.int64), index_ref.reverse_index_ref(wi))
        assert np.array_equal(p.get(_lib.PLAN_CU_WINDOW, np.int32), cu)
        assert np.array_equal(p.get(_lib.PLAN_CU_FULL, np.int32), index_ref.cu_seqlens_ref(grid))
        assert np.array_equal(p.get(_lib.PLAN_POS_IDS, np.int32).reshape(-1, 2), index_ref.rope_pos_ids_ref(grid))


def test_tower_7b_vs_hf_fp32_small_grid():
    """7B-shaped tower, real HF tower in fp32 on the GPU as the reference."""
    grid = [[2, 16, 16]]
    hf, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.float32, attn="sdpa", device=DEV)
    t = vit.B200VisionTower.from_hf(hf, device=DEV, return_dict=False)
    m = 512
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(7)).to(DEV)
    ref = hf_ref.hf_forward(hf, x, torch.tensor(grid, device=DEV))
    out = t(x, torch.tensor(grid))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_tower_7b_cfg2_frames_overlay_vs_hf_fp32():
    """BASELINE config 2: 16-frame 448x448 clip, box+mask overlay on all frames, bf16 CUDA path
    vs PIL-exact overlay -> HF processor layout -> HF tower fp32."""
    from PIL import Image, ImageDraw
    T, H, W = 16, 448, 448
    frames = hf_ref.synthetic_frames(T, H, W, clip_id=0)
    vip = Image.new("RGBA", (W, H), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    d.ellipse([(224 - 80, 224 - 80), (224 + 80, 224 + 80)], fill=(0, 255, 0, 100))
    d.rectangle([(112, 96), (335, 351)], outline=(255, 0, 0, 200), width=4)
    layer = np.array(vip)
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i - 8, sy=i - 8) for i in range(T)]
    ref_frames = overlay_ref.overlay_clip_ref(frames.numpy(), layer, [dict(mode=1, sx=i - 8, sy=i - 8) for i in range(T)])
    pv, grid = patchify_ref.patchify_ref(ref_frames)
    hf, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.float32, attn="sdpa", device=DEV)
    ref = hf_ref.hf_forward(hf, torch.from_numpy(pv).to(DEV), torch.tensor(grid, device=DEV))
    t = vit.B200VisionTower.from_hf(hf, device=DEV, return_dict=False)
    del hf
    out = t.forward_frames(frames.to(DEV), vit.OverlaySpec.from_rgba(layer, ops))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)
    # the pixel_values entry (HF boundary) must agree with the fused frames entry BIT FOR BIT: identical inputs to the
    # tower, and the balanced (stream-K) residual GEMMs add their partial tiles in a fixed order
    out2 = t(torch.from_numpy(pv).to(DEV), torch.tensor(grid))
    assert torch.equal(out2, out)
    assert torch.equal(t.forward_frames(frames.to(DEV), vit.OverlaySpec.from_rgba(layer, ops)), out)   # run to run
    # CUDA-graph replay; the returned tensor is the caller's own (a second replay must not overwrite it)
    t.use_cuda_graph = True
    out3 = t(torch.from_numpy(pv).to(DEV), torch.tensor(grid))
    out4 = t(torch.zeros_like(torch.from_numpy(pv)).to(DEV), torch.tensor(grid))
    assert torch.equal(out3, out) and not torch.equal(out4, out3)


@pytest.mark.parametrize("grid", [[[2, 48, 48]], [[16, 32, 32]]])
def test_tower_7b_cfg3_cfg4_shapes_vs_hf_fp32(grid):
    """The shapes BASELINE configs 3 and 4 are made of, against the real HF tower in fp32: one 672x672 temporal
    pair (2304-key full-attention slices, 48x48 patch grid = 36 windows per slice) and one 32-frame 448x448 clip."""
    hf, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=2, dtype=torch.float32, attn="sdpa", device=DEV)
    t = vit.B200VisionTower.from_hf(hf, device=DEV, return_dict=False)
    m = sum(a * b * c for a, b, c in grid)
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(9)).to(DEV)
    ref = hf_ref.hf_forward(hf, x, torch.tensor(grid, device=DEV))
    out = t(x, torch.tensor(grid))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_tower_7b_672_frames_overlay_vs_hf_fp32():
    """cfg 4 resolution through the fused entry: 4 frames of 672x672 with a box + scribble-like RGBA layer shifted per
    frame -> PIL-exact oracle overlay -> HF processor layout -> HF tower fp32."""
    from PIL import Image, ImageDraw
    T, H, W = 4, 672, 672
    frames = hf_ref.synthetic_frames(T, H, W, clip_id=4)
    vip = Image.new("RGBA", (W, H), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    d.rectangle([(150, 120), (520, 560)], outline=(255, 0, 0, 208), width=6)
    d.line([(100, 600), (300, 420), (420, 500), (600, 200)], fill=(0, 0, 255, 224), width=9)
    layer = np.array(vip)
    shifts = [(0, 0), (-13, 7), (25, -30), (3, 3)]
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=sx, sy=sy) for sx, sy in shifts]
    ref_frames = overlay_ref.overlay_clip_ref(frames.numpy(), layer, [dict(mode=1, sx=sx, sy=sy) for sx, sy in shifts])
    pv, grid = patchify_ref.patchify_ref(ref_frames)
    assert grid.tolist() == [[2, 48, 48]]
    hf, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=0, dtype=torch.float32, attn="sdpa", device=DEV)
    ref = hf_ref.hf_forward(hf, torch.from_numpy(pv).to(DEV), torch.tensor(grid, device=DEV))
    t = vit.B200VisionTower.from_hf(hf, device=DEV, return_dict=False)
    del hf
    out = t.forward_frames(frames.to(DEV), vit.OverlaySpec.from_rgba(layer, ops))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_overlay_layer_must_match_the_frames():
    """A prompt layer at another resolution (e.g. drawn before fit_frames resized the clip) raises, as
    Image.alpha_composite does in the reference, instead of reading out of bounds."""
    t, cfg, sd = make_tower(hf_ref.CFG_TINY)
    frames = hf_ref.synthetic_frames(2, 56, 84, clip_id=1).to(DEV)
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER)] * 2
    with pytest.raises(ValueError):
        t.forward_frames(frames, vit.OverlaySpec.from_rgba(np.zeros((112, 168, 4), np.uint8), ops))
    with pytest.raises(ValueError):
        t.forward_frames(frames, vit.OverlaySpec.from_palette(np.zeros((56, 85), np.uint8), [[0, 0, 0, 0]], ops))
    with pytest.raises(ValueError):
        t.forward_frames(frames, out=torch.empty(6, hf_ref.CFG_TINY["out_hidden_size"], dtype=torch.bfloat16))   # CPU out


def test_errors_are_loud():
    t, cfg, sd = make_tower(hf_ref.CFG_TINY)
    with pytest.raises(ValueError):
        t(torch.zeros(10, 1176, device=DEV), torch.tensor([[2, 8, 12]]))
    with pytest.raises(ValueError):
        t(torch.zeros(192, 1176), torch.tensor([[2, 8, 12]]))       # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        t.plan_for([[2, 7, 12]])                                     # odd h


# ------------------------------------------------------------------ BASELINE.json full sizes: size-independent properties
def _bench_tower():
    import bench
    t = vit.B200VisionTower(dict(hf_ref.CFG_7B), device=DEV, return_dict=False)
    bench.random_state_dict_gpu(t, seed=3)
    return t


def test_cfg4_long_video_slice_independence():
    """cfg 4: 64-frame 672x672 clip, grid_thw=[32,48,48] (73,728 patches, 2304-patch full-attention slices).
    No attention segment spans temporal slices (HF :488-496), so the clip's embeddings must equal the
    concatenation of its two 32-frame halves run separately."""
    t = _bench_tower()
    g = torch.Generator(device=DEV).manual_seed(5)
    frames = torch.randint(0, 256, (64, 672, 672, 3), dtype=torch.uint8, device=DEV, generator=g)
    whole = t.forward_frames(frames)
    assert whole.shape == (18432, 3584) and torch.isfinite(whole.float()).all()
    a = t.forward_frames(frames[:32])
    b = t.forward_frames(frames[32:])
    # not bit-equal: M differs, so the stream-K split points (which tiles add their k-range in two pieces) differ and
    # flip individual bf16 roundings downstream; the agreement stays at bf16-rounding level
    cos, rel = parity(torch.cat([a, b]), whole)
    assert cos >= 0.9999 and rel <= REL_MAX, (cos, rel)


def test_cfg3_batched_clips_equal_per_clip_runs():
    """cfg 3 shape: several [16,32,32] clips in ONE call (multi-grid plan, as HF concatenates videos) must equal
    per-clip calls, in clip order."""
    t = _bench_tower()
    n = 4
    g = torch.Generator(device=DEV).manual_seed(6)
    pv = torch.randn(n * 16384, 1176, device=DEV, generator=g).to(torch.bfloat16)
    whole = t(pv, torch.tensor([[16, 32, 32]] * n))
    assert whole.shape == (n * 4096, 3584)
    parts = [t(pv[i * 16384:(i + 1) * 16384], torch.tensor([[16, 32, 32]])) for i in range(n)]
    # same reasoning as above: the two schedules cut different tiles in two, so single bf16 roundings flip; both results
    # are bf16 renderings of the same fp32 function, so they agree within the path's own tolerance
    cos, rel = parity(torch.cat(parts), whole)
    assert cos >= 0.9999 and rel <= REL_MAX, (cos, rel)


def test_ragged_grid_7b_vs_hf_fp32():
    """Non-multiple-of-8 patch grids: short edge windows (HF get_window_index padding rule) at the 7B shape."""
    grid = [[2, 18, 14], [1, 6, 10]]
    hf, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_7B, seed=1, dtype=torch.float32, attn="sdpa", device=DEV)
    t = vit.B200VisionTower.from_hf(hf, device=DEV, return_dict=False)
    m = sum(a * b * c for a, b, c in grid)
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(8)).to(DEV)
    ref = hf_ref.hf_forward(hf, x, torch.tensor(grid, device=DEV))
    out = t(x, torch.tensor(grid))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)


def test_drop_in_inside_hf_qwen2_5_vl_model():
    """SURVEY.md 8b / cfg 5 boundary: swap the tower inside a full Qwen2.5-VL model (the class UniGRModel subclasses,
    /root/reference/model/qwen_2_5_vl_sam2.py:104) with `install()` and run the model's own forward: HF's
    get_video_features must accept our module (dtype/device attributes, .pooler_output) and the prefill logits must
    agree with the stock tower's."""
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    vc = dict(hf_ref.CFG_SMALL)
    vc["out_hidden_size"] = 256
    cfg = Qwen2_5_VLConfig(
        text_config=dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         intermediate_size=512, vocab_size=1000, max_position_embeddings=4096,
                         rope_scaling={"type": "mrope", "mrope_section": [8, 12, 12]}),
        vision_config=vc, video_token_id=990, image_token_id=991, vision_start_token_id=992, vision_end_token_id=993)
    torch.manual_seed(0)
    model = Qwen2_5_VLForConditionalGeneration(cfg).eval().to(DEV).to(torch.bfloat16)
    with torch.no_grad():  # non-trivial norm weights / biases in the tower
        for n, p_ in model.model.visual.named_parameters():
            if n.endswith("bias"):
                p_.normal_(0, 0.02)
            elif "norm" in n or "ln_q" in n:
                p_.copy_(1 + 0.1 * torch.randn_like(p_))
    grid = torch.tensor([[2, 8, 12], [1, 6, 10]], device=DEV)
    m = int(grid.prod(-1).sum())
    nvis = [int(g.prod()) // 4 for g in grid]
    ids = [1, 2]
    for n in nvis:
        ids += [992] + [990] * n + [993]
    ids = torch.tensor([ids + [5, 6, 7]], device=DEV)
    pv = torch.randn(m, 1176, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)).to(torch.bfloat16)
    with torch.no_grad():
        ref = model(input_ids=ids, pixel_values_videos=pv, video_grid_thw=grid).logits.float()
        tower = vit.install(model)
        assert isinstance(model.model.visual, vit.B200VisionTower) and tower.dtype == torch.bfloat16
        out = model(input_ids=ids, pixel_values_videos=pv, video_grid_thw=grid).logits.float()
    cos, rel = parity(out, ref)
    assert cos >= 0.999 and rel <= 2e-2, (cos, rel)   # both sides are bf16 end to end here


def test_splice_into_inputs_embeds_in_place():
    """the merger epilogue writes straight into the LLM's inputs_embeds rows; result equals HF's
    masked_scatter of the separately computed embeddings."""
    t, cfg, sd = make_tower(hf_ref.CFG_SMALL)
    grid = [[2, 8, 12]]
    m, n_vis, hid = 192, 48, hf_ref.CFG_SMALL["out_hidden_size"]
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(3)).to(DEV)
    ids = torch.tensor([[5, 6, 992] + [990] * n_vis + [993, 7, 8]], device=DEV)
    emb = torch.randn(1, ids.shape[1], hid, device=DEV, generator=torch.Generator(device=DEV).manual_seed(4)).to(torch.bfloat16)
    ref_tokens = t(x, torch.tensor(grid))
    mask = (ids == 990).unsqueeze(-1).expand_as(emb)
    ref = emb.clone().masked_scatter(mask, ref_tokens)                       # HF modeling :1309-1315
    start, length = vit.splice_span(ids, 990)
    assert (start, length) == (3, n_vis)
    t(x, torch.tensor(grid), out=emb[0, start:start + length])
    assert torch.equal(emb, ref)
    with pytest.raises(ValueError):
        vit.splice_span(torch.tensor([990, 1, 990]), 990)
    with pytest.raises(ValueError):
        t(x, torch.tensor(grid), out=emb[0, :length - 1])


def test_forward_frames_odd_frame_count_and_fp32_pixels():
    """Odd T (last frame repeated, HF videoproc :245-249) through the fused entry == oracle patchify -> tower; fp32 and
    fp16 pixel_values take the cast kernel."""
    t, cfg, sd = make_tower(hf_ref.CFG_SMALL, output_fp32=True)
    frames = hf_ref.synthetic_frames(5, 56, 84, clip_id=7)
    pv, grid = patchify_ref.patchify_ref(frames.numpy())
    ref = tower_ref.tower_forward_ref(sd, cfg, torch.from_numpy(pv), grid)
    out = t.forward_frames(frames.to(DEV))
    cos, rel = parity(out, ref)
    assert cos >= COS_MIN and rel <= REL_MAX, (cos, rel)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        o2 = t(torch.from_numpy(pv).to(DEV).to(dt), torch.from_numpy(grid))
        cos, rel = parity(o2, ref)
        assert cos >= COS_MIN and rel <= REL_MAX, (dt, cos, rel)


def test_clip_pipeline_matches_direct_calls():
    """ClipPipeline (overlapped H2D / compute / D2H, two clips in flight) returns, in order, exactly what
    forward_frames returns for each clip."""
    t, cfg, sd = make_tower(hf_ref.CFG_TINY)
    T, H, W = 4, 56, 84
    clips = [hf_ref.synthetic_frames(T, H, W, clip_id=i).pin_memory() for i in range(5)]
    layer = overlay_ref.box_layer_ref(H, W, (10, 8, 70, 48), 3, (255, 0, 0, 200))
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i, sy=-i) for i in range(T)]
    spec = vit.OverlaySpec.from_rgba(layer, ops)
    direct = [t.forward_frames(c.to(DEV), spec if i % 2 else None).clone() for i, c in enumerate(clips)]
    pipe = vit.ClipPipeline(t, (T, H, W, 3), depth=2)
    got = []
    for i, c in enumerate(clips):
        r = pipe.submit(c, spec if i % 2 else None)
        if r is not None:
            got.append(r.clone())
    got += [r.clone() for r in pipe.drain()]
    assert len(got) == len(clips)
    for a, b in zip(got, direct):
        assert torch.equal(a, b.cpu())
    assert pipe.h2d_bytes == T * H * W * 3 and pipe.d2h_bytes == direct[0].numel() * direct[0].element_size()
    with pytest.raises(ValueError):
        vit.ClipPipeline(t, (T, H, W, 3), depth=1)


def test_shared_workspace_and_plan_lru_across_many_grids():
    """One workspace per slot (grown to the largest grid) and a bounded plan cache: alternating resolutions gives the
    same embeddings as a fresh tower per grid, and memory does not pile up per grid."""
    t, cfg, sd = make_tower(hf_ref.CFG_TINY, max_plans=3)
    grids = [[[1, 4, 6]], [[2, 8, 12]], [[1, 2, 2]], [[3, 6, 10]], [[1, 4, 6]], [[2, 8, 12]], [[1, 10, 8]], [[1, 2, 2]]]
    outs = []
    for i, g in enumerate(grids):
        m = sum(a * b * c for a, b, c in g)
        x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(100 + i)).to(DEV)
        outs.append((g, x, t(x, g).clone()))
        assert len(t._plans) <= 3 and len(t._workspaces) == 1
    ws_bytes = t._workspaces[0][2]
    assert ws_bytes == max(vit.B200VisionTower(dict(hf_ref.CFG_TINY), device="cpu").plan_for(g).ws_bytes for g in grids)
    for g, x, out in outs:
        fresh, _, _ = make_tower(hf_ref.CFG_TINY)
        assert torch.equal(fresh(x, g), out), g


def test_module_apply_keeps_the_tower_consistent():
    """`.to(dtype)` / `.to(device)` on the nn.Module re-packs the weights instead of running on stale copies."""
    t, cfg, sd = make_tower(hf_ref.CFG_TINY)
    g = [[2, 4, 6]]
    x = torch.randn(48, 1176, generator=torch.Generator().manual_seed(3)).to(DEV)
    ref = t(x, g).clone()
    t.to(torch.float32)                       # parameters become fp32 (same values): same packed bf16 weights
    assert t._packed is None and t.dtype == torch.float32
    out32 = t(x, g)                           # the output dtype follows the parameters, as on the HF module
    assert out32.dtype == torch.float32 and torch.equal(out32.to(torch.bfloat16), ref)
    ref = out32.clone()
    with torch.no_grad():
        t.blocks[0].mlp.down_proj.bias.add_(1.0)
    assert torch.equal(t(x, g), ref)          # an in-place edit is invisible until the pack is rebuilt ...
    t.invalidate()                            # ... which invalidate() (or any _apply / load_state_dict) asks for
    assert not torch.equal(t(x, g), ref)
    t.half()
    assert t.dtype == torch.float16 and t(x, g).dtype == torch.bfloat16   # dtype follows the parameters; outputs stay bf16
