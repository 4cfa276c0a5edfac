"""Mask / scribble prompt layers on the B200 (csrc/raster.cu through the C ABI): coverage bit-equal to the oracle, to
Pillow itself at full size, and to the layers the reference's own draw_mask / draw_scribble produced; composited frames
equal to the reference's alpha_composite."""
import os

import numpy as np
import pytest
import torch
from PIL import Image, ImageDraw

pytestmark = pytest.mark.gpu

import rga3_release_b200 as vit
from rga3_release_b200 import _lib
from oracle import raster_ref
from test_raster_cpu import contour_polygon, pil_line, pil_polygon, random_polygon

DEV = "cuda"


def test_polygons_match_oracle_and_pillow():
    rng = np.random.default_rng(3)
    for t in range(400):
        w, h = int(rng.integers(6, 64)), int(rng.integers(6, 64))
        xy = random_polygon(rng, w, h, t % 3)
        got = vit.mask_layer([[v for p in xy for v in p]], h, w, DEV).cpu().numpy()
        want = np.zeros((h, w), np.uint8)
        raster_ref.fill_polygon_ref(want, xy)
        assert np.array_equal(got, want), (w, h, xy)
    for t in range(300):                                       # full-size contours and random polygons vs live Pillow
        w, h = int(rng.integers(100, 700)), int(rng.integers(100, 700))
        xy = contour_polygon(rng, w, h) if t % 2 else random_polygon(rng, w, h, t % 3)
        got = vit.mask_layer([[v for p in xy for v in p]], h, w, DEV).cpu().numpy()
        assert np.array_equal(got, pil_polygon(w, h, xy)), (w, h)


def test_lines_match_pillow():
    rng = np.random.default_rng(4)
    for t in range(600):
        w, h = int(rng.integers(10, 500)), int(rng.integers(10, 500))
        width = int(rng.integers(1, 60)) if t % 5 else 1
        p0 = (float(rng.uniform(-5, w + 5)), float(rng.uniform(-5, h + 5)))
        p1 = (p0[0] + float(rng.normal(0, 3)), p0[1] + float(rng.normal(0, 3))) if t % 2 else \
             (float(rng.uniform(-5, w + 5)), float(rng.uniform(-5, h + 5)))
        got = vit.lines_layer([p0, p1], width, h, w, DEV).cpu().numpy()
        assert np.array_equal(got, pil_line(w, h, p0, p1, width)), (w, h, p0, p1, width)


def test_reference_golden_layers(golden_dir):
    """Layers drawn by /root/reference/utils/visual_prompt_generator.py itself (tests/golden/make_prompt_golden.py)."""
    z = np.load(os.path.join(golden_dir, "prompt_layers.npz"))
    for ci in range(4):
        h, w = (int(v) for v in z[f"mask{ci}_hw"])
        segs = [z[f"mask{ci}_seg{s}"].tolist() for s in range(int(z[f"mask{ci}_nseg"]))]
        want = np.unpackbits(z[f"mask{ci}_cov"])[: h * w].reshape(h, w)
        got = vit.mask_layer(segs, h, w, DEV)
        assert np.array_equal(got.cpu().numpy(), want), ci
        assert vit.get_bbox_from_mask(got) == tuple(z[f"mask{ci}_bbox"].tolist())
    for ci in range(6):
        h, w, anchor, width = (int(v) for v in z[f"scribble{ci}_params"])
        want = np.unpackbits(z[f"scribble{ci}_cov"])[: h * w].reshape(h, w)
        got = vit.scribble_layer(z[f"scribble{ci}_ctrl"], width, h, w, DEV, image_size_anchor=anchor).cpu().numpy()
        assert np.array_equal(got, want), ci


def test_image_blending_scribble_end_to_end(golden_dir):
    """image_blending(shape='scribble') of the reference -> blended RGB frame; here: prompt_line_width + from_scribble +
    the overlay kernel.  Bit-equal frame."""
    z = np.load(os.path.join(golden_dir, "prompt_layers.npz"))
    h, w, anchor, width, r, g, b, alpha = (int(v) for v in z["blend_params"])
    lw = vit.prompt_line_width("scribble", w, h, anchor, width=width)
    spec = vit.OverlaySpec.from_scribble(z["blend_ctrl"], (r, g, b), alpha, lw, h, w, [vit.FrameOp(mode=_lib.FRAME_LAYER)],
                                         device=DEV, image_size_anchor=anchor)
    want_cov = np.unpackbits(z["blend_layer_cov"])[: h * w].reshape(h, w)
    assert np.array_equal(spec.layer.cpu().numpy(), want_cov)
    frames = torch.from_numpy(z["blend_frame"]).unsqueeze(0).to(DEV)
    out = torch.empty_like(frames)
    fr = _lib.Frames(frames.data_ptr(), 1, h, w)
    _lib.check(_lib.lib().b200vit_overlay_composite(fr, spec.to_c(1), out.data_ptr(), torch.cuda.current_stream().cuda_stream), "composite")
    assert np.array_equal(out[0].cpu().numpy(), z["blend_out"])


def test_mask_prompt_composite_equals_pillow_pipeline():
    """shape='mask' the way image_blending builds it (:294-299, :352-363): contours filled with rgb + (alpha,), layer
    alpha-composited onto the frame -- against Pillow doing exactly that."""
    rng = np.random.default_rng(5)
    h, w = 448, 448
    segs = [[int(v) for p in contour_polygon(rng, w, h) for v in p] for _ in range(2)]
    rgb, alpha = (0, 255, 0), 100
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    vip = Image.new("RGBA", (w, h), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    for seg in segs:
        d.polygon([(seg[i], seg[i + 1]) for i in range(0, len(seg), 2)], outline=None, fill=rgb + (alpha,), width=1)
    want = np.array(Image.alpha_composite(Image.fromarray(frame).convert("RGBA"), vip).convert("RGB"))
    spec = vit.OverlaySpec.from_mask(segs, rgb, alpha, h, w, [vit.FrameOp(mode=_lib.FRAME_LAYER)], device=DEV)
    frames = torch.from_numpy(frame).unsqueeze(0).to(DEV)
    out = torch.empty_like(frames)
    _lib.check(_lib.lib().b200vit_overlay_composite(_lib.Frames(frames.data_ptr(), 1, h, w), spec.to_c(1), out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "composite")
    assert np.array_equal(out[0].cpu().numpy(), want)


def test_raster_rejects_bad_input():
    with pytest.raises(ValueError):
        vit.mask_layer([[0, 0, 1e9, 5, 3, 3]], 32, 32, DEV)          # coordinate out of range
    with pytest.raises(ValueError):
        vit.lines_layer([(0, 0), (float("nan"), 3)], 4, 32, 32, DEV)
    assert int(vit.mask_layer([], 16, 16, DEV).sum()) == 0
