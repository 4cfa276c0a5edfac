"""Golden prompt layers drawn by the REFERENCE's own code: /root/reference/utils/visual_prompt_generator.py is imported
unmodified (its unavailable third-party imports -- pycocotools, skimage, matplotlib, shapely -- are stubbed; none of
them is touched by the functions used here) and draw_mask (:268-274), draw_scribble (:230-252), get_bbox_from_mask
(:406-414) and image_blending (:284-368, scribble without segmentation) are called with fixed inputs.  Run in the build
container (needs /root/reference and Pillow):  python tests/golden/make_prompt_golden.py
Writes tests/golden/prompt_layers.npz (coverage bit-packed)."""
import os
import random
import sys
import types

import numpy as np
from PIL import Image, ImageDraw

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    for name in ("pycocotools", "pycocotools.mask", "skimage", "skimage.measure", "matplotlib", "matplotlib.pyplot",
                 "shapely", "shapely.ops", "shapely.geometry", "shapely.validation"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["shapely.ops"].unary_union = None
    sys.modules["shapely.geometry"].Point = None
    sys.modules["shapely.geometry"].Polygon = None
    sys.modules["shapely.validation"].explain_validity = None
    sys.path.insert(0, REF)
    import utils.visual_prompt_generator as vpg
    return vpg


def blob_contour(h, w, seed):
    """A closed integer contour the way mask_to_segmentation_coords emits them: the rounded boundary of a smooth blob,
    duplicates removed, ring closed (flat [x0, y0, x1, y1, ...])."""
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(0.3, 0.7) * w, rng.uniform(0.3, 0.7) * h
    k = rng.integers(2, 5)
    amp = rng.uniform(0.05, 0.3)
    r0 = rng.uniform(0.12, 0.3) * min(h, w)
    ang = np.linspace(0, 2 * np.pi, 1200, endpoint=False)
    r = r0 * (1 + amp * np.sin(k * ang + rng.uniform(0, 6.28)))
    pts = np.round(np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], 1)).astype(int)
    coords = list(dict.fromkeys(map(tuple, pts.tolist())))
    coords.append(coords[0])
    return [int(v) for p in coords for v in p]


def main():
    vpg = import_reference()
    out = {}
    cases = []
    # ---- masks: draw_mask on (h, w) with one or more contours
    for ci, (h, w, n_seg) in enumerate([(448, 448, 1), (336, 448, 2), (672, 672, 3), (120, 90, 1)]):
        segs = [blob_contour(h, w, 100 * ci + s) for s in range(n_seg)]
        img = Image.new("RGBA", (w, h), (0, 0, 0, 0))
        vpg.draw_mask(ImageDraw.Draw(img), None, segs, (0, 255, 0, 100), 1)
        cov = (np.array(img)[..., 3] > 0)
        out[f"mask{ci}_cov"] = np.packbits(cov)
        out[f"mask{ci}_hw"] = np.array([h, w])
        out[f"mask{ci}_nseg"] = np.array(n_seg)
        for s, seg in enumerate(segs):
            out[f"mask{ci}_seg{s}"] = np.array(seg, dtype=np.int64)
        out[f"mask{ci}_bbox"] = np.array(vpg.get_bbox_from_mask(cov), dtype=np.int64)
        cases.append(("mask", ci))
    # ---- scribbles: draw_scribble with the control points the reference would sample, captured by replacing its sampler
    for ci, (h, w, anchor, width, seed) in enumerate([(448, 448, 448, 13, 1), (336, 448, 448, 12, 2), (672, 672, 448, 21, 3),
                                                       (448, 448, 336, 17, 4), (100, 140, 448, 3, 5), (60, 60, 448, 1, 6)]):
        rng = np.random.default_rng(seed)
        ctrl = [(float(rng.uniform(-0.05, 1.05) * w), float(rng.uniform(-0.05, 1.05) * h)) for _ in range(4)]
        it = iter(ctrl)
        vpg.get_random_point_within_bbox = lambda bbox: next(it)
        img = Image.new("RGBA", (w, h), (0, 0, 0, 0))
        vpg.draw_scribble(ImageDraw.Draw(img), (0, 0, w, h), None, (255, 0, 0, 200), width, max_image_size=max(w, h),
                          image_size_anchor=anchor)
        out[f"scribble{ci}_cov"] = np.packbits(np.array(img)[..., 3] > 0)
        out[f"scribble{ci}_params"] = np.array([h, w, anchor, width])
        out[f"scribble{ci}_ctrl"] = np.array(ctrl, dtype=np.float64)
        cases.append(("scribble", ci))
    # ---- image_blending end to end (scribble, no segmentation): frame in, blended frame + prompt layer out
    h, w = 224, 308
    rng = np.random.default_rng(9)
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ctrl = [(float(rng.uniform(0, w)), float(rng.uniform(0, h))) for _ in range(4)]
    it = iter(ctrl)
    vpg.get_random_point_within_bbox = lambda bbox: next(it)
    random.seed(5)
    blended, vip = vpg.image_blending(Image.fromarray(frame), shape="scribble", bbox_coord=(0, 0, w, h), segmentation=None,
                                      image_size_anchor=448, rgb_value=(0, 0, 255), alpha=210, width=14, return_vip_img=True)
    out["blend_frame"] = frame
    out["blend_out"] = np.array(blended)
    out["blend_layer_cov"] = np.packbits(np.array(vip)[..., 3] > 0)
    out["blend_ctrl"] = np.array(ctrl, dtype=np.float64)
    out["blend_params"] = np.array([h, w, 448, 14, 0, 0, 255, 210])
    np.savez_compressed(os.path.join(HERE, "prompt_layers.npz"), **out)
    print("wrote", len(out), "arrays;", cases)


if __name__ == "__main__":
    main()
