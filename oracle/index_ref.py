"""Integer index oracles for the vision tower (numpy, loop-level restatement).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

HF = transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py (5.5.0), the code the
reference reaches through model/qwen_2_5_vl_sam2.py:182-200.
"""
from __future__ import annotations

import numpy as np


def window_index_ref(grid_thw, window_size=112, spatial_merge_size=2, patch_size=14):
    """HF:modeling:411-451 ``get_window_index`` followed by the
    ``unique_consecutive`` of HF:modeling:476.

    Returns (window_index int64 [sum t*h*w/4], cu_window_seqlens_raw list,
    cu_window_seqlens int32 after unique_consecutive).
    """
    win = window_size // spatial_merge_size // patch_size
    unit = spatial_merge_size * spatial_merge_size
    window_index = []
    cu = [0]
    base = 0
    for t, h, w in np.asarray(grid_thw, dtype=np.int64).tolist():
        lh, lw = h // spatial_merge_size, w // spatial_merge_size
        # pad = win - n % win: a whole extra (empty) window when n % win == 0 (:424-425)
        pad_h = win - lh % win
        pad_w = win - lw % win
        nwh = (lh + pad_h) // win
        nww = (lw + pad_w) // win
        for ti in range(t):
            for wh in range(nwh):
                for ww in range(nww):
                    cnt = 0
                    for ih in range(win):
                        for iw in range(win):
                            y = wh * win + ih
                            x = ww * win + iw
                            if y < lh and x < lw:
                                window_index.append(base + ti * lh * lw + y * lw + x)
                                cnt += 1
                    cu.append(cu[-1] + cnt * unit)
        base += t * lh * lw
    raw = list(cu)
    uniq = [raw[0]]
    for v in raw[1:]:
        if v != uniq[-1]:
            uniq.append(v)
    return (np.asarray(window_index, dtype=np.int64), raw, np.asarray(uniq, dtype=np.int32))


def reverse_index_ref(window_index):
    """HF:modeling:512 ``torch.argsort(window_index)`` (window_index is a permutation)."""
    rev = np.empty_like(window_index)
    rev[window_index] = np.arange(window_index.shape[0], dtype=window_index.dtype)
    return rev


def cu_seqlens_ref(grid_thw):
    """HF:modeling:488-496: one segment per temporal slice of h*w patches."""
    out = [0]
    for t, h, w in np.asarray(grid_thw, dtype=np.int64).tolist():
        for _ in range(t):
            out.append(out[-1] + h * w)
    return np.asarray(out, dtype=np.int32)


def rope_pos_ids_ref(grid_thw, spatial_merge_size=2):
    """HF:modeling:382-404: (hpos, wpos) per patch in merge-group-major order
    (h/2, w/2, mh, mw), repeated over t.  Returns int64 [M, 2]."""
    m = spatial_merge_size
    out = []
    for t, h, w in np.asarray(grid_thw, dtype=np.int64).tolist():
        one = []
        for bh in range(h // m):
            for bw in range(w // m):
                for ih in range(m):
                    for iw in range(m):
                        one.append((bh * m + ih, bw * m + iw))
        out.extend(one * t)
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


def rope_table_ref(grid_thw, head_dim=80, theta=10000.0, spatial_merge_size=2):
    """HF:modeling:117-130 + :405-409: rotary angles [M, head_dim/2] in fp32,
    in ORIGINAL (un-reordered) patch order.  cat(hpos*inv_freq, wpos*inv_freq)."""
    dim = head_dim // 2
    import torch  # torch's fp32 pow, so the table is bit-identical to HF's buffer
    inv_freq = (1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))).numpy()
    pos = rope_pos_ids_ref(grid_thw, spatial_merge_size)
    fh = pos[:, 0:1].astype(np.float32) * inv_freq[None, :]
    fw = pos[:, 1:2].astype(np.float32) * inv_freq[None, :]
    return np.concatenate([fh, fw], axis=1).astype(np.float32)


def reorder_rows_ref(x, window_index, unit=4):
    """HF:modeling:478-481: x.reshape(M/4,4,-1)[window_index].reshape(M,-1)."""
    m = x.shape[0]
    return x.reshape(m // unit, unit, -1)[window_index].reshape(m, -1)
