"""The prompt-layer oracle (oracle/raster_ref.py) against Pillow itself and against the layers the reference's own
draw_mask / draw_scribble produced (tests/golden/prompt_layers.npz); host-side helpers of prompts.py."""
import os
import random

import numpy as np
import pytest
from PIL import Image, ImageDraw

import rga3_release_b200 as vit
from oracle import raster_ref


def pil_polygon(w, h, xy):
    img = Image.new("RGBA", (w, h), (0, 0, 0, 0))
    ImageDraw.Draw(img).polygon(xy, outline=None, fill=(255, 0, 0, 100), width=1)       # draw_mask's call (:274)
    return (np.array(img)[..., 3] > 0).astype(np.uint8)


def pil_line(w, h, p0, p1, width):
    img = Image.new("RGBA", (w, h), (0, 0, 0, 0))
    ImageDraw.Draw(img).line([p0, p1], fill=(255, 0, 0, 200), width=width)               # draw_scribble's call (:250)
    return (np.array(img)[..., 3] > 0).astype(np.uint8)


def random_polygon(rng, w, h, kind):
    n = int(rng.integers(3, 12))
    if kind == 0:        # integer vertices, some outside the image
        return [(int(rng.integers(-5, w + 5)), int(rng.integers(-5, h + 5))) for _ in range(n)]
    if kind == 1:        # float vertices (truncated by Pillow)
        return [(float(rng.uniform(-5, w + 5)), float(rng.uniform(-5, h + 5))) for _ in range(n)]
    xy, x, y = [], int(rng.integers(0, w)), int(rng.integers(0, h))   # rectilinear: runs of horizontal / vertical sides
    for _ in range(n):
        if rng.random() < 0.5:
            x = int(rng.integers(0, w))
        else:
            y = int(rng.integers(0, h))
        xy.append((x, y))
    return xy


def contour_polygon(rng, w, h):
    """unit-step closed contour (what mask_to_segmentation_coords emits, :372-403)"""
    ang = np.linspace(0, 2 * np.pi, 400, endpoint=False)
    r = rng.uniform(0.15, 0.4) * min(w, h) * (1 + rng.uniform(0, 0.4) * np.sin(rng.integers(2, 6) * ang + rng.uniform(0, 6)))
    pts = np.round(np.stack([w / 2 + r * np.cos(ang), h / 2 + r * np.sin(ang)], 1)).astype(int)
    c = list(dict.fromkeys(map(tuple, pts.tolist())))
    return c + [c[0]]


def test_polygon_fill_oracle_matches_pillow():
    rng = np.random.default_rng(0)
    for t in range(1500):
        w, h = int(rng.integers(6, 64)), int(rng.integers(6, 64))
        xy = random_polygon(rng, w, h, t % 3)
        im = np.zeros((h, w), np.uint8)
        raster_ref.fill_polygon_ref(im, xy)
        assert np.array_equal(im, pil_polygon(w, h, xy)), (w, h, xy)
    for t in range(40):
        w, h = int(rng.integers(40, 120)), int(rng.integers(40, 120))
        xy = contour_polygon(rng, w, h)
        im = np.zeros((h, w), np.uint8)
        raster_ref.fill_polygon_ref(im, xy)
        assert np.array_equal(im, pil_polygon(w, h, xy)), (w, h)


def test_line_oracle_matches_pillow():
    rng = np.random.default_rng(1)
    for t in range(3000):
        w, h = int(rng.integers(10, 80)), int(rng.integers(10, 80))
        width = int(rng.integers(1, 40)) if t % 5 else 1
        p0 = (float(rng.uniform(-5, w + 5)), float(rng.uniform(-5, h + 5)))
        p1 = (p0[0] + float(rng.normal(0, 3)), p0[1] + float(rng.normal(0, 3))) if t % 2 else \
             (float(rng.uniform(-5, w + 5)), float(rng.uniform(-5, h + 5)))
        im = np.zeros((h, w), np.uint8)
        raster_ref.line_ref(im, p0, p1, width)
        assert np.array_equal(im, pil_line(w, h, p0, p1, width)), (w, h, p0, p1, width)


def test_oracle_matches_reference_golden_layers(golden_dir):
    z = np.load(os.path.join(golden_dir, "prompt_layers.npz"))
    for ci in (3,):                                           # the small mask case (pure-Python oracle: seconds)
        h, w = z[f"mask{ci}_hw"]
        segs = [z[f"mask{ci}_seg{s}"].tolist() for s in range(int(z[f"mask{ci}_nseg"]))]
        want = np.unpackbits(z[f"mask{ci}_cov"])[: h * w].reshape(h, w)
        assert np.array_equal(raster_ref.mask_layer_ref(segs, h, w), want)
        assert vit.get_bbox_from_mask(want) == tuple(z[f"mask{ci}_bbox"].tolist())
    for ci in (4, 5):                                         # small scribbles (width 3 and the Bresenham width 1)
        h, w, anchor, width = z[f"scribble{ci}_params"]
        want = np.unpackbits(z[f"scribble{ci}_cov"])[: h * w].reshape(h, w)
        got = raster_ref.scribble_layer_ref(z[f"scribble{ci}_ctrl"].tolist(), int(width), int(h), int(w), int(anchor))
        assert np.array_equal(got, want), ci


def test_scribble_points_match_the_reference_expression():
    rng = np.random.default_rng(2)
    for _ in range(20):
        ctrl = rng.uniform(-50, 700, (4, 2))
        w, h, anchor = int(rng.integers(60, 900)), int(rng.integers(60, 900)), int(rng.choice([336, 448]))
        got = vit.scribble_points(ctrl, w, h, anchor)
        want = raster_ref.scribble_points_ref(ctrl.tolist(), int(1000 * max(w, h) / anchor))
        assert got.shape == want.shape and np.array_equal(got, want)


def test_prompt_size_rules():
    """image_blending :294-297 / :326-359 restated in prompt_alpha / prompt_line_width."""
    for seed in range(50):
        w, h = random.Random(seed).randint(100, 1300), random.Random(seed + 1).randint(100, 1300)
        s = max(w, h) / 448
        r = random.Random(seed)
        assert 188 <= vit.prompt_alpha("scribble", r) <= 224 and 72 <= vit.prompt_alpha("mask", r) <= 128
        lw = vit.prompt_line_width("scribble", w, h, 448, rng=random.Random(seed))
        assert lw == max(random.Random(seed).randint(int(12 * s), int(15 * s)), 1)
        assert vit.prompt_line_width("rectangle", w, h, 448, visual_prompt_style="constant") == max(int(3 * s), 1)
        assert vit.prompt_line_width("mask", w, h, 448, width=2) == max(int(2 * s), 1)
        assert vit.prompt_line_width("mask", w, h, 448, rng=random.Random(seed)) == random.Random(seed).randint(0, int(2 * s))
    with pytest.raises(ValueError):
        vit.prompt_line_width("hexagon", 10, 10)
    m = np.zeros((20, 30), bool)
    m[3:7, 10:25] = True
    assert vit.get_bbox_from_mask(m) == (10, 3, 25, 7)
    with pytest.raises(IndexError):
        vit.get_bbox_from_mask(np.zeros((4, 4), bool))
