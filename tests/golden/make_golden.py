"""Generate the golden fixtures in this directory from the REFERENCE itself.

Run in the build container only (needs /root/reference, transformers, Pillow,
cv2):   python tests/golden/make_golden.py
The fixtures are committed; tests and the GPU box never read /root/reference.

What is run:
  * /root/reference/model/STOM.py  ``STOM.warp`` / ``STOM.warp_point`` -- imported
    unmodified with ``cotracker`` stubbed in sys.modules (the tracker model is
    out of scope; only the overlay functions are called).
  * Pillow ``Image.alpha_composite`` / ``ImageDraw.rectangle`` the way
    /root/reference/utils/visual_prompt_generator.py:102-104, :361-363 call them
    (that module itself needs shapely/skimage/pycocotools, absent here).
  * HF ``Qwen2VLVideoProcessor`` and ``Qwen2_5_VisionTransformerPretrainedModel``
    (the third-party code behind /root/reference/model/qwen_2_5_vl_sam2.py:182-200).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference_stom():
    for name in ("cotracker", "cotracker.utils", "cotracker.utils.visualizer", "cotracker.predictor"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["cotracker.utils.visualizer"].Visualizer = object
    sys.modules["cotracker.predictor"].CoTrackerPredictor = object
    sys.path.insert(0, REF)
    from model.STOM import STOM  # noqa
    return STOM


def gen_overlay():
    from PIL import Image, ImageDraw
    STOM = import_reference_stom()
    stom = STOM.__new__(STOM)  # no tracker checkpoint; warp/warp_point do not touch self
    rng = np.random.default_rng(7)
    h, w = 56, 84
    frames = rng.integers(0, 256, (6, h, w, 3), dtype=np.uint8)
    # prompt layer drawn like image_blending: box outline + filled disc ("mask")
    vip = Image.new("RGBA", (w, h), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    d.ellipse([(30, 14), (60, 44)], fill=(0, 255, 0, 100))
    d.rectangle([(10, 8), (70, 48)], outline=(255, 0, 0, 200), width=3)
    layer = np.array(vip)
    flows = np.array([[0.0, 0.0], [3.0, -2.0], [-4.7, 5.2], [-0.7, -0.4], [40.5, 10.25], [-90.0, 3.0]], dtype=np.float32)
    warped, comps = [], []
    for i in range(6):
        # reference call convention: warp(vip, frame, avg_flow_y(=x flow), avg_flow_x(=y flow)), STOM.py:133-135
        pil, wv = stom.warp(layer.copy(), frames[i].copy(), flows[i, 0], flows[i, 1])
        warped.append(np.array(wv))
        comps.append(np.array(pil))
    key = np.array(Image.alpha_composite(Image.fromarray(frames[0], "RGB").convert("RGBA"), vip).convert("RGB"))
    # warp_point: visible tracks around a centre -> circle stamp
    tracks = np.stack([rng.uniform(30, 50, 40), rng.uniform(15, 40, 40)], axis=1).astype(np.float32)  # (x, y)
    vis = np.ones(40, dtype=bool)
    pil, wv = stom.warp_point(layer.copy(), frames[1].copy(), tracks, vis)
    wp_layer = np.array(wv)
    wp_comp = np.array(pil)
    np.savez_compressed(os.path.join(HERE, "overlay_stom.npz"), frames=frames, layer=layer, flows=flows,
                        warped=np.stack(warped), comps=np.stack(comps), key=key,
                        wp_tracks=tracks, wp_layer=wp_layer, wp_comp=wp_comp)
    # rectangle KATs incl. degenerate widths
    boxes = np.array([[10, 12, 40, 50, 3], [0, 0, 63, 63, 5], [5, 5, 8, 8, 4], [20, 10, 20, 30, 1], [-4, 3, 30, 70, 2],
                      [13, 4, 14, 13, 5], [12, 42, 18, 54, 11]], dtype=np.int64)
    outs = []
    for l, t, r, b, wd in boxes.tolist():
        img = Image.new("RGBA", (64, 64), (0, 0, 0, 0))
        ImageDraw.Draw(img).rectangle([(l, t), (r, b)], outline=(0, 0, 255, 190), width=wd)
        outs.append(np.array(img))
    np.savez_compressed(os.path.join(HERE, "overlay_boxes.npz"), boxes=boxes, layers=np.stack(outs))



def policy_case(rng, h, w, t, n, key_idx, shape):
    """Synthetic tracker output around a prompt: a common per-frame motion + jitter, outliers, frames with few
    visible points, one frame with none, one with a non-finite visible coordinate."""
    from PIL import Image, ImageDraw
    yy, xx = np.mgrid[0:h, 0:w]
    frames = np.stack([((xx * 3 + yy * 2 + 17 * i) % 256).astype(np.uint8)[..., None].repeat(3, 2) for i in range(t)])
    frames[..., 1] = 255 - frames[..., 1]
    vip = Image.new("RGBA", (w, h), (0, 0, 0, 0))
    d = ImageDraw.Draw(vip)
    if shape == "mask":
        d.ellipse([(w // 3, h // 4), (w // 3 + w // 4, h // 4 + h // 3)], fill=(0, 255, 0, 160))
    else:
        d.rectangle([(w // 4, h // 5), (w // 4 + w // 3, h // 5 + h // 2)], outline=(255, 0, 0, 200), width=3)
    layer = np.array(vip)
    base = np.stack([rng.uniform(w * 0.3, w * 0.6, n), rng.uniform(h * 0.25, h * 0.6, n)], axis=1)
    tracks = np.zeros((t, n, 2), dtype=np.float32)
    vis = np.ones((t, n), dtype=bool)
    for i in range(t):
        motion = np.array([(i - key_idx) * 2.3 - 0.4 * (i % 3), (key_idx - i) * 1.7 + 0.3])
        tracks[i] = (base + motion + rng.normal(0, 0.4, (n, 2))).astype(np.float32)
        out = rng.random(n) < 0.1
        tracks[i][out] += rng.normal(0, 25, (int(out.sum()), 2)).astype(np.float32)
        vis[i] = rng.random(n) < 0.85
    tracks[key_idx] = base.astype(np.float32)
    if t > 3:
        vis[(key_idx + 1) % t] = rng.random(n) < 0.3      # too few visible -> untouched
        vis[(key_idx + 2) % t] = False                     # none visible
    if t > 5:
        j = (key_idx + 3) % t
        k = int(np.nonzero(vis[j])[0][0])
        tracks[j, k, 0] = np.nan                           # visible NaN: mask shapes raise -> untouched; others -> NaN mean
    if t > 6:
        tracks[(key_idx + 4) % t] += np.float32(-0.6)      # a flow that truncates toward zero at the border
    return frames, layer, tracks, vis


def gen_policy():
    """STOM.propagate_in_video (the reference's own function, tracker stubbed) on synthetic tracks."""
    from PIL import Image
    STOM = import_reference_stom()
    stom = STOM.__new__(STOM)
    rng = np.random.default_rng(11)
    out = {}
    cases = [("rect_a", 75, 90, 8, 37, 2, "rectangle"), ("rect_b", 84, 56, 7, 200, 0, "scribble"),
             ("mask_odd", 150, 135, 8, 60, 3, "mask"), ("mask_even", 150, 180, 7, 45, 6, "mask contour"),
             ("mask_small", 60, 90, 5, 9, 1, "mask")]
    for name, h, w, t, n, key_idx, shape in cases:
        frames, layer, tracks, vis = policy_case(rng, h, w, t, n, key_idx, shape)
        stom.track_in_video = lambda fr, vip, idx, save_path, tr=tracks, vi=vis: (tr[None].copy(), vi[None].copy())
        pil_frames = [Image.fromarray(f, "RGB") for f in frames]
        res = stom.propagate_in_video(pil_frames, Image.fromarray(layer, "RGBA"), key_idx, shape=shape)
        out[name + "_frames"] = frames
        out[name + "_layer"] = layer
        out[name + "_tracks"] = tracks
        out[name + "_vis"] = vis
        out[name + "_meta"] = np.array([key_idx, 1 if shape in ("mask", "mask contour") else 0], dtype=np.int64)
        out[name + "_out"] = np.stack([np.array(im.convert("RGB")) for im in res])
    np.savez_compressed(os.path.join(HERE, "stom_policy.npz"), **out)


def gen_resize():
    """PIL.Image.resize (default BICUBIC) -- the call qwen_vl_utils.fetch_image makes per frame -- and HF smart_resize."""
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize
    rng = np.random.default_rng(21)
    out = {}
    cases = [(45, 80, 28, 56), (36, 64, 56, 84), (90, 60, 112, 84), (61, 47, 28, 28), (72, 128, 72, 84), (50, 100, 84, 100)]
    for i, (h, w, oh, ow) in enumerate(cases):
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([(xx * 5 + yy * 3) % 256, (xx * yy) % 256, rng.integers(0, 256, (h, w))], axis=2).astype(np.uint8)
        out[f"in{i}"] = img
        out[f"out{i}"] = np.array(Image.fromarray(img, "RGB").resize((ow, oh)))
    sizes = []
    for _ in range(400):
        h, w = int(rng.integers(20, 2400)), int(rng.integers(20, 2400))
        mp = int(rng.choice([384 * 28 * 28, 336 * 28 * 28, 1280 * 28 * 28, 16384 * 28 * 28]))
        mn = int(rng.choice([4 * 28 * 28, 56 * 56, 256 * 28 * 28]))
        if max(h, w) / min(h, w) > 200:
            continue
        oh, ow = smart_resize(h, w, factor=28, min_pixels=mn, max_pixels=mp)
        sizes.append((h, w, mn, mp, oh, ow))
    out["smart_resize"] = np.asarray(sizes, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "resize_pil.npz"), **out)

def gen_patchify():
    from oracle import hf_ref
    fr = hf_ref.synthetic_frames(3, 56, 84, clip_id=3)   # odd T: exercises the pad-by-repeat
    pv, grid = hf_ref.hf_patchify(fr)
    np.savez_compressed(os.path.join(HERE, "patchify_hf.npz"), frames=fr.numpy(), pixel_values=pv.numpy(),
                        grid_thw=grid.numpy())


def gen_index():
    from oracle import hf_ref
    model, _, _ = hf_ref.build_hf_tower(hf_ref.CFG_TINY)
    out = {}
    for i, grid in enumerate([[[2, 8, 12]], [[8, 32, 32]], [[1, 6, 10], [2, 18, 14]], [[1, 2, 2]], [[2, 48, 48]]]):
        g = torch.tensor(grid)
        wi, cu = model.get_window_index(g)
        out[f"grid{i}"] = np.asarray(grid)
        out[f"window_index{i}"] = wi.numpy()
        out[f"cu_window_raw{i}"] = np.asarray(cu)
        out[f"rot{i}"] = model.rot_pos_emb(g).numpy()
    np.savez_compressed(os.path.join(HERE, "index_hf.npz"), **out)


def gen_tower():
    from oracle import hf_ref
    model, cfg, sd = hf_ref.build_hf_tower(hf_ref.CFG_TINY, seed=0)
    grid = [[1, 6, 10], [2, 8, 12]]
    m = sum(t * h * w for t, h, w in grid)
    x = torch.randn(m, 1176, generator=torch.Generator().manual_seed(11))
    out = hf_ref.hf_forward(model, x, torch.tensor(grid))
    np.savez_compressed(os.path.join(HERE, "tower_tiny.npz"), grid_thw=np.asarray(grid), x_seed=11,
                        pooler_output=out.numpy().astype(np.float32))


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only == "policy":
        gen_policy()
        sys.exit(0)
    if only == "resize":
        gen_resize()
        sys.exit(0)
    gen_overlay()
    gen_policy()
    gen_resize()
    gen_patchify()
    gen_index()
    gen_tower()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
