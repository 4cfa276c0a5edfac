"""STOM placement policy on the B200 (csrc/stom_policy.cu) through the C ABI: device frame ops against the host policy
and the oracle, and the composited frames (ops never leaving the device) against the golden frames of the reference's
own ``STOM.propagate_in_video``."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import rga3_release_b200 as vit
from rga3_release_b200 import _lib
from oracle import stom_policy_ref as sp

DEV = "cuda"
CASES = ["rect_a", "rect_b", "mask_odd", "mask_even", "mask_small"]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "stom_policy.npz"))
    fr, layer, tr, vi, meta, out = [g[f"{name}_{k}"] for k in ("frames", "layer", "tracks", "vis", "meta", "out")]
    return fr, layer, tr, vi, int(meta[0]), ("mask" if meta[1] else "rectangle"), out


def _device_ops(tr, vi, key, shape, h, w, layer):
    lay = torch.from_numpy(layer).to(DEV) if layer is not None else None
    ops, r = vit.stom_frame_ops_device(torch.from_numpy(tr).to(DEV), torch.from_numpy(vi).to(DEV), key, shape, h, w, lay)
    return ops, r, lay


def _as_tuples(ops):
    out = []
    for o in ops:
        if o.mode == _lib.FRAME_LAYER:
            out.append((1, o.sx, o.sy, int(bool(o.zx)), int(bool(o.zy))))
        elif o.mode == _lib.FRAME_CIRCLE:
            out.append((2, o.cx, o.cy, o.r, tuple(int(v) for v in o.rgba)))
        else:
            out.append((0,))
    return out


@pytest.mark.parametrize("name", CASES)
def test_device_policy_golden_frames(golden_dir, name):
    fr, layer, tr, vi, key, shape, out = _case(golden_dir, name)
    t, h, w = fr.shape[:3]
    dops, r, lay = _device_ops(tr, vi, key, shape, h, w, layer)
    # same ops as the host policy (numpy + cv2)
    host = vit.stom_frame_ops(tr, vi, key, shape, h, w, layer)
    assert _as_tuples(vit.frame_ops_from_bytes(dops.cpu().numpy())) == _as_tuples(host)
    # and the composited clip, with the ops consumed straight from device memory, is the reference's output
    spec = vit.OverlaySpec(kind=_lib.LAYER_RGBA, layer=lay, device_ops=dops, device_ops_circle_r=r)
    frd = torch.from_numpy(fr).to(DEV)
    fc = _lib.Frames(frd.data_ptr(), t, h, w)
    comp = torch.zeros_like(frd)
    _lib.check(_lib.lib().b200vit_overlay_composite(C.byref(fc), C.byref(spec.to_c(t)), comp.data_ptr(), _stream()), "composite")
    torch.cuda.synchronize()
    assert np.array_equal(comp.cpu().numpy(), out)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (8, 3), (129, 4), (1000, 5), (4097, 6), (16384, 7)])
def test_device_flow_policy_random(n, seed):
    """Flow branch on random tracks of every size class of numpy's pairwise summation, incl. ties and outliers."""
    rng = np.random.default_rng(seed)
    t, h, w, key = 6, 448, 448, 2
    base = rng.uniform(50, 400, (n, 2))
    tr = np.zeros((t, n, 2), dtype=np.float32)
    vi = rng.random((t, n)) < 0.9
    for i in range(t):
        tr[i] = (base + rng.normal(0, 3, (1, 2)) * (i - key) + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
        out = rng.random(n) < 0.08
        tr[i][out] += rng.normal(0, 40, (int(out.sum()), 2)).astype(np.float32)
    tr[key] = base.astype(np.float32)
    tr[4] = np.round(tr[4])                       # many identical magnitudes -> ties around the median
    vi[5] = rng.random(n) < 0.45                   # borderline visibility
    dops, r, _ = _device_ops(tr, vi, key, "rectangle", h, w, None)
    got = _as_tuples(vit.frame_ops_from_bytes(dops.cpu().numpy()))
    want = []
    for i in range(t):
        if i == key:
            want.append((1, 0, 0, 0, 0))
            continue
        fl = sp.flow_shift_ref(tr[key], tr[i], vi[i])
        if fl is None:
            want.append((0,))
        else:
            sx, zx = vit.shift_from_flow(float(fl[0]), w)
            sy, zy = vit.shift_from_flow(float(fl[1]), h)
            want.append((1, sx, sy, zx, zy))
    assert got == want
    assert r == -1


@pytest.mark.parametrize("h,w,n,seed", [(448, 448, 300, 0), (672, 448, 50, 1), (150, 180, 2000, 2), (28, 56, 5, 3)])
def test_device_point_policy_random(h, w, n, seed):
    """Mask branch (warp_point) at real frame sizes: closing with the 29 / 44-wide ellipse, centroid, colour."""
    rng = np.random.default_rng(seed)
    t, key = 5, 0
    layer = np.zeros((h, w, 4), dtype=np.uint8)
    layer[h // 3: h // 2, w // 4: w // 2] = (10, 200, 30, 40 + 60 * seed)   # alpha below / inside / above the clamp range
    tr = np.zeros((t, n, 2), dtype=np.float32)
    vi = rng.random((t, n)) < 0.8
    for i in range(t):
        c = np.array([rng.uniform(0.2, 0.8) * w, rng.uniform(0.2, 0.8) * h])
        tr[i] = (c + rng.normal(0, min(h, w) / 10, (n, 2))).astype(np.float32)   # some land outside the frame
    vi[3] = rng.random(n) < 0.3
    tr[4, :, 0] -= w                                                             # everything out of bounds -> m00 == 0
    dops, r, _ = _device_ops(tr, vi, key, "mask", h, w, layer)
    got = _as_tuples(vit.frame_ops_from_bytes(dops.cpu().numpy()))
    want = []
    for i in range(t):
        if i == key:
            want.append((1, 0, 0, 0, 0))
            continue
        st = sp.point_stamp_ref(layer, tr[i], vi[i])
        want.append((0,) if st is None or st == "empty" else (2, st[0], st[1], st[2], tuple(st[3])))
    assert got == want
    assert r == min(h, w) // 20


def test_device_ops_through_forward_frames():
    """forward_frames with device-resident ops == forward_frames with the same ops given on the host."""
    from oracle import hf_ref, tower_ref
    cfg = tower_ref.TowerCfg(**hf_ref.CFG_TINY)
    tower = vit.B200VisionTower(dict(hf_ref.CFG_TINY), device=DEV, return_dict=False)
    tower.load_state_dict(hf_ref.make_state_dict(cfg, 0))
    t, h, w, n, key = 4, 56, 84, 40, 1
    rng = np.random.default_rng(9)
    frames = torch.from_numpy(rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)).to(DEV)
    layer = np.zeros((h, w, 4), dtype=np.uint8)
    layer[10:30, 20:50] = (255, 0, 0, 200)
    base = rng.uniform(20, 50, (n, 2))
    tr = np.stack([(base + (i - key) * np.array([2.6, -1.3]) + rng.normal(0, 0.2, (n, 2))) for i in range(t)]).astype(np.float32)
    vi = np.ones((t, n), dtype=bool)
    grid = torch.tensor([[t // 2, h // 14, w // 14]])
    host_ops = vit.stom_frame_ops(tr, vi, key, "rectangle", h, w, layer)
    out_host = tower.forward_frames(frames, vit.OverlaySpec.from_rgba(layer, host_ops), grid_thw=grid)
    dops, r, lay = _device_ops(tr, vi, key, "rectangle", h, w, layer)
    out_dev = tower.forward_frames(frames, vit.OverlaySpec(kind=_lib.LAYER_RGBA, layer=lay, device_ops=dops,
                                                          device_ops_circle_r=r), grid_thw=grid)
    assert any(o.mode == _lib.FRAME_LAYER and (o.sx or o.sy) for o in host_ops)
    assert torch.equal(out_host, out_dev)
    plain = tower.forward_frames(frames, None, grid_thw=grid)
    assert not torch.equal(plain, out_dev)        # the overlay did change the embeddings


def test_device_policy_rejects_bad_args():
    l = _lib.lib()
    ops = torch.zeros((4, 36), dtype=torch.uint8, device=DEV)
    trk = torch.zeros((4, 8, 2), dtype=torch.float32, device=DEV)
    vis = torch.ones((4, 8), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)
    args = lambda **kw: [kw.get("trk", trk.data_ptr()), vis.data_ptr(), 4, kw.get("n", 8), kw.get("key", 0), kw.get("mask", 0),
                         kw.get("h", 56), 84, kw.get("layer", None), ops.data_ptr(), ws.data_ptr(), kw.get("wsb", ws.numel()), _stream()]
    assert l.b200vit_stom_policy(*args(key=4)) == -1                 # key frame out of range
    assert l.b200vit_stom_policy(*args(n=16385)) == -1               # too many points
    assert l.b200vit_stom_policy(*args(wsb=16)) == -1                # workspace too small
    assert l.b200vit_stom_policy(*args(mask=1)) == -1                # mask shapes need the layer
    assert l.b200vit_stom_policy(*args(mask=1, h=10, layer=ws.data_ptr())) == -1   # min(h,w)/15 == 0
    assert l.b200vit_stom_policy(*args()) == 0
    torch.cuda.synchronize()
