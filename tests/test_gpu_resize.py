"""Bicubic resize on the B200 (csrc/resize.cu) through the C ABI: bit-exact against the oracle, the PIL golden fixture
and (when Pillow is importable on the box) live ``Image.resize``; plus the fit -> overlay -> tower chain."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import rga3_release_b200 as vit
from rga3_release_b200 import _lib
from oracle import resize_ref as rr

DEV = "cuda"


def test_resize_golden_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "resize_pil.npz"))
    i = 0
    while f"in{i}" in g:
        img, ref = g[f"in{i}"], g[f"out{i}"]
        out = vit.resize_frames(torch.from_numpy(img)[None].to(DEV), ref.shape[0], ref.shape[1])
        assert np.array_equal(out[0].cpu().numpy(), ref), i
        i += 1


@pytest.mark.parametrize("t,h,w,oh,ow", [(3, 72, 128, 56, 84), (2, 360, 640, 252, 448), (1, 90, 60, 112, 84),
                                         (2, 100, 56, 28, 56), (2, 56, 100, 56, 28), (1, 64, 64, 64, 64),
                                         (1, 1080, 1920, 448, 784)])
def test_resize_matches_oracle_and_pil(t, h, w, oh, ow):
    rng = np.random.default_rng(h * 7 + w)
    frames = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    out = vit.resize_frames(torch.from_numpy(frames).to(DEV), oh, ow).cpu().numpy()
    try:
        from PIL import Image
        ref = np.stack([np.array(Image.fromarray(f, "RGB").resize((ow, oh))) for f in frames])
    except ImportError:
        ref = np.stack([rr.resize_bicubic_ref(f, oh, ow) for f in frames])
    assert np.array_equal(out, ref)
    if h * w <= 360 * 640:
        assert np.array_equal(out[0], rr.resize_bicubic_ref(frames[0], oh, ow))


def test_fit_then_forward_frames():
    """A clip that is not a multiple of 28: fit_frames (smart_resize + resize) feeds forward_frames; same result as
    resizing with the oracle on the host."""
    from oracle import hf_ref, tower_ref
    cfg = tower_ref.TowerCfg(**hf_ref.CFG_TINY)
    tower = vit.B200VisionTower(dict(hf_ref.CFG_TINY), device=DEV, return_dict=False)
    tower.load_state_dict(hf_ref.make_state_dict(cfg, 0))
    rng = np.random.default_rng(4)
    frames = rng.integers(0, 256, (4, 75, 101, 3), dtype=np.uint8)
    oh, ow = vit.smart_resize(75, 101, 28, 4 * 28 * 28, 12 * 28 * 28)
    assert (oh % 28, ow % 28) == (0, 0) and oh * ow <= 12 * 28 * 28
    fitted = vit.fit_frames(torch.from_numpy(frames).to(DEV), 4 * 28 * 28, 12 * 28 * 28)
    assert tuple(fitted.shape) == (4, oh, ow, 3)
    host = np.stack([rr.resize_bicubic_ref(f, oh, ow) for f in frames])
    assert np.array_equal(fitted.cpu().numpy(), host)
    a = tower.forward_frames(fitted)
    b = tower.forward_frames(torch.from_numpy(host).to(DEV))
    assert torch.equal(a, b)


def test_resize_rejects_bad_args():
    l = _lib.lib()
    x = torch.zeros((1, 30, 40, 3), dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):
        vit.resize_frames(x, 28, 30)                       # width not a multiple of 4
    with pytest.raises(ValueError):
        vit.resize_frames(x.cpu(), 28, 28)
    with pytest.raises(ValueError):
        vit.resize_frames(x, 28, 28, out=torch.zeros((1, 28, 56, 3), dtype=torch.uint8, device=DEV))
    assert l.b200vit_resize_bicubic(x.data_ptr(), 1, 30, 40, x.data_ptr(), 28, 28, None, 0,
                                    torch.cuda.current_stream().cuda_stream) == -1
