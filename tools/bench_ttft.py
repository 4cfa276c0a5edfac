"""BASELINE config 5: UniGR-7B-shaped prefill (random-init Qwen2.5-VL-7B: 28-layer LLM + 32-layer vision tower) on one
cfg-2 clip + 32 text tokens; time-to-first-token with the stock HF tower (bf16, flash_attention_2 as the reference
loads it, /root/reference/app.py:50-56), with the B200 tower installed, and with the B200 tower writing its tokens
straight into inputs_embeds + host-built M-RoPE ids (splice_span / mrope_position_ids).
The class: /root/reference does not exist on the GPU box, and UniGRModel.forward(past_key_values=...) is exactly
`super().forward` of this HF class (qwen_2_5_vl_sam2.py:143-146; the SAM2 head is not on the prefill path) --
tests/test_gpu_unigr.py pins that equivalence on logits the reference class itself produced."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rga3_release_b200 as vit
from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration

DEV = "cuda"
vc = dict(depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16, out_hidden_size=3584, window_size=112,
          fullatt_block_indexes=[7, 15, 23, 31])
tc = dict(hidden_size=3584, num_hidden_layers=28, num_attention_heads=28, num_key_value_heads=4, intermediate_size=18944,
          vocab_size=152064, max_position_embeddings=32768, rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]})
res = {}
for attn in ("flash_attention_2", "sdpa"):
    try:
        cfg = Qwen2_5_VLConfig(text_config=tc, vision_config=vc, video_token_id=151656, image_token_id=151655,
                               vision_start_token_id=151652, vision_end_token_id=151653)
        cfg._attn_implementation = attn
        torch.manual_seed(0)
        with torch.device(DEV):
            model = Qwen2_5_VLForConditionalGeneration(cfg).to(torch.bfloat16).eval()
        break
    except Exception as ex:
        print(attn, "unavailable:", repr(ex)[:200])
res["llm_attn"] = attn
grid = torch.tensor([[8, 32, 32]], device=DEV)
ids = torch.tensor([[1] * 16 + [151652] + [151656] * 2048 + [151653] + [2] * 14], device=DEV)
pv = torch.randn(8192, 1176, device=DEV).to(torch.bfloat16)
types = (ids == 151656).to(torch.int32) * 2          # mm_token_type_ids: HF builds the M-RoPE ids from them (get_rope_index)


def ttft(n=5):
    ts = []
    with torch.no_grad():
        for i in range(n + 2):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = model(input_ids=ids, pixel_values_videos=pv, video_grid_thw=grid, mm_token_type_ids=types)
            tok = out.logits[:, -1].argmax(-1)
            e.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


res["ttft_ms_hf_tower"] = ttft()
tower = vit.install(model)
res["ttft_ms_b200_tower"] = ttft()
res["speedup"] = res["ttft_ms_hf_tower"] / res["ttft_ms_b200_tower"]

# splice: the merger epilogue writes the visual tokens into the LLM's inputs_embeds rows; position ids from mrope.py
start, n_vis = vit.splice_span(ids, 151656)
embed = model.get_input_embeddings()


def ttft_splice(n=5):
    ts = []
    with torch.no_grad():
        for i in range(n + 2):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            emb = embed(ids)
            tower(pv, grid, out=emb[0, start:start + n_vis])
            pos, _ = vit.mrope_position_ids(ids, types, video_grid_thw=grid, tokens_per_second=cfg.vision_config.tokens_per_second)
            out = model(inputs_embeds=emb, position_ids=pos)
            tok = out.logits[:, -1].argmax(-1)
            e.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2], out.logits[:, -1].float()


with torch.no_grad():
    ref_last = model(input_ids=ids, pixel_values_videos=pv, video_grid_thw=grid, mm_token_type_ids=types).logits[:, -1].float()
res["ttft_ms_b200_tower_spliced"], last = ttft_splice()
res["spliced_last_logits_max_abs_diff"] = float((last - ref_last).abs().max())
res["speedup_spliced"] = res["ttft_ms_hf_tower"] / res["ttft_ms_b200_tower_spliced"]
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ttft_cfg5.json", "w"), indent=1)
