"""Rescale + normalise + 2x14x14 patchify oracle (numpy fp32).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

HF = transformers 5.5.0: image_processing_backends.py:291-331
(``rescale_and_normalize``), models/qwen2_vl/video_processing_qwen2_vl.py:239-272
(patchify), utils/constants.py (OPENAI_CLIP_MEAN/STD).  The reference calls this
through ``processor(text, videos=...)`` at /root/reference/utils/dataset.py:77-84
and evaluation/videoinfer/inference_videoinfer.py:301-308.
"""
from __future__ import annotations

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def normalise_lut_ref():
    """fp32 table lut[c, v] = (float(v) - mean_c*255) / (std_c*255), with the
    operation order of the fused HF path (mean*(1/rescale) in fp32 first)."""
    inv = np.float32(1.0 / (1.0 / 255.0))
    mean = np.asarray(CLIP_MEAN, dtype=np.float32) * inv
    std = np.asarray(CLIP_STD, dtype=np.float32) * inv
    v = np.arange(256, dtype=np.float32)[None, :]
    return ((v - mean[:, None]) / std[:, None]).astype(np.float32)


def patchify_ref(frames_u8: np.ndarray, patch=14, tps=2, merge=2):
    """frames [T, H, W, 3] uint8 -> (pixel_values fp32 [t*h*w, 3*tps*patch*patch],
    grid_thw).  Rows ordered (t, h/merge, w/merge, mh, mw); columns
    (c, tp, ph, pw).  Odd T is padded by repeating the last frame
    (video_processing_qwen2_vl.py:245-249)."""
    t_in, hh, ww, c = frames_u8.shape
    assert c == 3 and hh % (patch * merge) == 0 and ww % (patch * merge) == 0
    lut = normalise_lut_ref()
    if t_in % tps:
        frames_u8 = np.concatenate([frames_u8, np.repeat(frames_u8[-1:], tps - t_in % tps, axis=0)], axis=0)
    tt = frames_u8.shape[0]
    gt, gh, gw = tt // tps, hh // patch, ww // patch
    x = np.empty((tt, 3, hh, ww), dtype=np.float32)
    for ch in range(3):
        x[:, ch] = lut[ch][frames_u8[..., ch]]
    x = x.reshape(gt, tps, 3, gh // merge, merge, patch, gw // merge, merge, patch)
    #            0   1   2   3           4      5      6           7      8
    x = x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return (np.ascontiguousarray(x).reshape(gt * gh * gw, 3 * tps * patch * patch),
            np.asarray([[gt, gh, gw]], dtype=np.int64))
