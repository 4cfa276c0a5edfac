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
    for h, w, mn, mp, oh, ow in rows.tolist():
        assert vit.smart_resize(h, w, 28, mn, mp) == (oh, ow)
        assert rr.smart_resize_ref(h, w, 28, mn, mp) == (oh, ow)
    with pytest.raises(ValueError):
        vit.smart_resize(10, 4000)
