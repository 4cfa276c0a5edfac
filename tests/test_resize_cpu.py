"""Resize step (smart_resize + Pillow bicubic) on the CPU: oracle and host arithmetic against the golden fixture
(generated from PIL / HF by tests/golden/make_golden.py) and against the live libraries."""
import os

import numpy as np
import pytest

from oracle import resize_ref as rr


def test_oracle_matches_pil_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "resize_pil.npz"))
    i = 0
    while f"in{i}" in g:
        ref = g[f"out{i}"]
        assert np.array_equal(rr.resize_bicubic_ref(g[f"in{i}"], ref.shape[0], ref.shape[1]), ref), i
        i += 1
    assert i >= 6


def test_oracle_matches_live_pil():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    for h, w, oh, ow in [(72, 128, 56, 84), (240, 426, 196, 336), (90, 60, 112, 84), (37, 53, 28, 28), (64, 64, 64, 64),
                         (50, 80, 140, 196), (300, 40, 28, 28)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img, "RGB").resize((ow, oh)))
        assert np.array_equal(rr.resize_bicubic_ref(img, oh, ow), ref), (h, w, oh, ow)


def test_smart_resize_host_and_oracle(golden_dir):
    import rga3_release_b200 as vit
    rows = np.load(os.path.join(golden_dir, "resize_pil.npz"))["smart_resize"]
    assert len(rows) > 300
    differ = 0
    for h, w, mn, mp, oh, ow in rows.tolist():
        assert vit.smart_resize(h, w, 28, mn, mp, variant="transformers") == (oh, ow)     # fixture = HF's function
        assert rr.smart_resize_ref(h, w, 28, mn, mp) == (oh, ow)
        q = vit.smart_resize(h, w, 28, mn, mp)                                            # default: qwen_vl_utils 0.0.10
        differ += q != (oh, ow)
        if min(h, w) >= 14 and (oh * ow <= mp):
            assert q == (oh, ow), (h, w, mn, mp)        # the two sources agree away from tiny sides / the max_pixels floor
    # known answers of the qwen_vl_utils 0.0.10 rule (what the reference calls, utils/dataset.py:76): the first
    # rounding is floored at `factor`, so a 10-px side does not push the frame into the min_pixels branch
    assert vit.smart_resize(10, 1000) == (28, 1008) and vit.smart_resize(10, 1000, variant="transformers") == (28, 560)
    assert vit.smart_resize(448, 448) == (448, 448) and vit.smart_resize(720, 1280, 28, 3136, 320 * 28 * 28) == (364, 644)
    assert vit.smart_resize(13, 13) == (56, 56)
    with pytest.raises(ValueError):
        vit.smart_resize(10, 4000)
    with pytest.raises(ValueError):
        vit.smart_resize(10, 10, variant="pil")
