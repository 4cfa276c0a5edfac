"""Token-by-token restatement of the two published M-RoPE index rules (test infrastructure).

``rope_index_449_ref`` follows the Qwen2.5-VL release the reference pins (transformers 4.49.0.dev0,
/root/reference/requirements.txt:25; ``Qwen2_5_VLForConditionalGeneration.get_rope_index``): that version is not
installed in this image, so this function is a restatement of the published algorithm and parity for it is UNPINNED.
``rope_index_55_ref`` follows transformers 5.5.0 modeling_qwen2_5_vl.py:968-1133 and is checked against the installed
function in tests/test_mrope_cpu.py.  Plain Python loops, small inputs only."""
from __future__ import annotations


def _vision_runs(types):
    runs, i = [], 0
    while i < len(types):
        j = i
        while j < len(types) and types[j] == types[i]:
            j += 1
        runs.append((types[i], i, j))
        i = j
    return runs


def rope_index_449_ref(types, image_grids, video_grids, second_per_grid_ts, merge=2, tokens_per_second=2):
    """types: per-token 0/1/2 of ONE unpadded sequence.  Returns 3 lists (t, h, w positions)."""
    img, vid, spg = iter(image_grids or []), iter(video_grids or []), iter(second_per_grid_ts or [])
    out = [[], [], []]
    nxt = 0                                   # st_idx: max position so far + 1
    for kind, s, e in _vision_runs(list(types)):
        if kind == 0:
            for k in range(e - s):
                for a in range(3):
                    out[a].append(nxt + k)
            nxt += e - s
            continue
        t, h, w = next(img) if kind == 1 else next(vid)
        sec = 0.0 if kind == 1 else (next(spg) if second_per_grid_ts is not None else 1.0)
        gh, gw = h // merge, w // merge
        base, top = nxt, nxt
        for ti in range(t):
            tp = int(ti * sec * tokens_per_second)          # time_tensor.long(): truncation per token
            for hi in range(gh):
                for wi in range(gw):
                    out[0].append(base + tp)
                    out[1].append(base + hi)
                    out[2].append(base + wi)
                    top = max(top, base + tp, base + hi, base + wi)
        nxt = top + 1
    return out


def rope_index_55_ref(types, image_grids, video_grids, second_per_grid_ts, merge=2, tokens_per_second=2):
    img, vid = iter(image_grids or []), iter(video_grids or [])
    spg = iter(second_per_grid_ts) if second_per_grid_ts is not None else None
    out = [[], [], []]
    cur = 0
    for kind, s, e in _vision_runs(list(types)):
        if kind == 0:
            for k in range(e - s):
                for a in range(3):
                    out[a].append(cur + k)
            cur += e - s
            continue
        t, h, w = next(img) if kind == 1 else next(vid)
        interval = tokens_per_second * int(next(spg) if spg is not None else 1)
        gh, gw = h // merge, w // merge
        n = gh * gw * t
        for i in range(n):
            out[0].append(cur * interval)
            out[1].append(cur + i // (gw * t))                # arange(h).repeat_interleave(w * t)
            out[2].append(cur + i % gw)                       # arange(w).repeat(h * t)
        cur += max(h, w) // merge
    return out
