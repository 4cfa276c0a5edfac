"""CPU-only checks of the host logic: the C ABI loads and exports every declared symbol,
plan construction (pure host code) matches the oracle bit for bit, the overlay policy
reduces to the reference's per-pixel loop, multi-rank sharding/gather works (gloo, world 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import rga3_release_b200 as vit
from rga3_release_b200 import _lib
from oracle import hf_ref, index_ref, overlay_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200vit.h")).read()
    declared = set(re.findall(r"\b(b200vit_[a-z0-9_]+)\s*\(", hdr))
    lib = vit.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in b200vit.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.b200vit_version() == _lib.VERSION == 3


@pytest.mark.parametrize("grid", [[[2, 8, 12]], [[8, 32, 32]], [[1, 6, 10], [2, 18, 14]], [[1, 2, 2]], [[3, 48, 48]]])
def test_plan_host_arrays_match_oracle(grid):
    t = vit.B200VisionTower(dict(hf_ref.CFG_TINY), device="cpu")
    p = t.plan_for(grid)
    wi, raw, cu = index_ref.window_index_ref(grid)
    assert np.array_equal(p.get(_lib.PLAN_WINDOW_INDEX, np.int64), wi)
    rev = index_ref.reverse_index_ref(wi)
    assert np.array_equal(p.get(_lib.PLAN_REVERSE_INDEX, np.int64), rev)
    assert np.array_equal(p.get(_lib.PLAN_CU_WINDOW, np.int32), cu)
    assert np.array_equal(p.get(_lib.PLAN_CU_FULL, np.int32), index_ref.cu_seqlens_ref(grid))
    assert np.array_equal(p.get(_lib.PLAN_POS_IDS, np.int32).reshape(-1, 2), index_ref.rope_pos_ids_ref(grid))
    m = p.m
    row_map = p.get(_lib.PLAN_ROW_MAP, np.int32)
    # x.reshape(M/4,4,-1)[window_index] (HF :478-481) as a scatter: row r lands at row_map[r]
    x = np.arange(m, dtype=np.int64)[:, None]
    scat = np.empty_like(x)
    scat[row_map] = x
    assert np.array_equal(scat, index_ref.reorder_rows_ref(x, wi))
    # rope tables: window-ordered cos/sin of the oracle's angles (fp32 libm vs torch: 1e-6)
    ang = index_ref.reorder_rows_ref(index_ref.rope_table_ref(grid), wi)
    cos = p.get(_lib.PLAN_ROPE_COS, np.float32).reshape(m, 40)
    sin = p.get(_lib.PLAN_ROPE_SIN, np.float32).reshape(m, 40)
    assert np.abs(cos - np.cos(ang)).max() < 2e-6 and np.abs(sin - np.sin(ang)).max() < 2e-6
    # what the QKV epilogue reads: HF's table by coordinate (fp32, as HF rotates, :149-167) + window-ordered positions;
    # gathering one with the other reproduces the [M, 40] tables above bit for bit
    tab = p.get(_lib.PLAN_ROPE_TABLE, np.float32).reshape(-1, 20, 2)
    pos = p.get(_lib.PLAN_ROPE_POS, np.int32).reshape(m, 2)
    assert tab.shape[0] == max(max(g[1], g[2]) for g in grid)
    assert np.array_equal(pos, index_ref.reorder_rows_ref(index_ref.rope_pos_ids_ref(grid), wi))
    gathered = np.concatenate([tab[pos[:, 0]], tab[pos[:, 1]]], axis=1)
    assert np.array_equal(gathered[..., 0], cos) and np.array_equal(gathered[..., 1], sin)


def test_plan_rejects_bad_grids():
    t = vit.B200VisionTower(dict(hf_ref.CFG_TINY), device="cpu")
    for bad in ([[2, 7, 12]], [[0, 8, 8]], [[1, 8, -2]]):
        with pytest.raises(ValueError):
            t.plan_for(bad)
    with pytest.raises(ValueError):
        vit.B200VisionTower(dict(hf_ref.CFG_TINY, hidden_size=128, num_heads=2), device="cpu").plan_for([[1, 4, 4]])  # head_dim 64


def test_state_dict_layout_matches_hf():
    cfg = dict(hf_ref.CFG_TINY)
    t = vit.B200VisionTower(cfg, device="cpu")
    hf, _, sd = hf_ref.build_hf_tower(cfg)
    ours = {k: tuple(v.shape) for k, v in t.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in hf.state_dict().items() if "inv_freq" not in k}
    assert ours == theirs
    t.load_state_dict(sd)
    assert torch.equal(t.blocks[1].attn.qkv.bias.float(), sd["blocks.1.attn.qkv.bias"].to(torch.bfloat16).float())


def test_shift_from_flow_matches_reference_loop():
    rng = np.random.default_rng(11)
    for _ in range(40):
        lay = np.zeros((14, 17, 4), np.uint8)
        m = rng.random((14, 17)) < 0.5
        lay[m] = rng.integers(1, 256, (int(m.sum()), 4), dtype=np.uint8)
        fx, fy = np.float32(rng.uniform(-5, 5)), np.float32(rng.uniform(-5, 5))
        sx, zx = vit.shift_from_flow(fx, 17)
        sy, zy = vit.shift_from_flow(fy, 14)
        assert (sx, zx) == overlay_ref.shift_params_ref(fx, 17) and (sy, zy) == overlay_ref.shift_params_ref(fy, 14)
        assert np.array_equal(overlay_ref.warp_layer_ref(lay, fx, fy), overlay_ref.warp_layer_gather_ref(lay, sx, zx, sy, zy))


def test_stom_frame_ops_policy(golden_dir):
    """Policy against the reference's own outputs: translated frames of the golden STOM.warp run and the
    warp_point circle stamp."""
    z = np.load(os.path.join(golden_dir, "overlay_stom.npz"))
    layer, flows = z["layer"], z["flows"]
    h, w = layer.shape[:2]
    n = 30
    key = np.stack([np.linspace(20, 60, n), np.linspace(10, 40, n)], 1).astype(np.float32)
    tracks = np.stack([key + flows[i] for i in range(len(flows))]).astype(np.float32)
    vis = np.ones((len(flows), n), bool)
    ops = vit.stom_frame_ops(tracks, vis, 0, "rectangle", h, w, layer)
    for i in range(len(flows)):
        assert ops[i].mode == _lib.FRAME_LAYER
        lay = overlay_ref.warp_layer_gather_ref(layer, ops[i].sx, ops[i].zx, ops[i].sy, ops[i].zy)
        assert np.array_equal(lay, z["warped"][i]), i
    # mask shape -> circle stamp identical to the reference's warp_point layer
    tr = np.stack([z["wp_tracks"], z["wp_tracks"]])
    ops = vit.stom_frame_ops(tr, np.ones((2, tr.shape[1]), bool), 0, "mask", h, w, layer)
    o = ops[1]
    assert o.mode == _lib.FRAME_CIRCLE
    assert np.array_equal(overlay_ref.circle_layer_ref(h, w, o.cx, o.cy, o.r, o.rgba), z["wp_layer"])
    # too few visible points -> frame left untouched (STOM.py:165-166, :122)
    ops = vit.stom_frame_ops(tr, np.zeros((2, tr.shape[1]), bool), 0, "mask", h, w, layer)
    assert ops[1].mode == _lib.FRAME_NONE


def test_shard_helpers():
    assert [vit.shard_clips(8, 4, r) for r in range(4)] == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert [vit.shard_clips(5, 4, r) for r in range(4)] == [[0, 1], [2], [3], [4]]
    assert [vit.shard_slices(32, 8, r) for r in (0, 7)] == [(0, 4), (28, 32)]
    assert sum(len(vit.shard_clips(3, 8, r)) for r in range(8)) == 3


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["B200_ROOT"])
import rga3_release_b200 as vit
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
clips = vit.shard_clips(5, world, rank)
rows = [len(vit.shard_clips(5, world, r)) * 3 for r in range(world)]
local = torch.cat([torch.full((3, 4), float(c)) for c in clips]) if clips else torch.empty(0, 4)
out = vit.gather_tokens(local, rows, dst=0)
if rank == 0:
    ref = torch.cat([torch.full((3, 4), float(c)) for c in range(5)])
    assert torch.equal(out, ref), out
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
"""


def test_gather_tokens_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, B200_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def build_c_demo(tmp_path):
    """gcc -std=c99 -pedantic -Werror on examples/c_abi_demo.c against libb200vit.so (+ libcudart for memory calls)."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    vit.lib()  # make sure the library exists (raises with the build hint otherwise)
    libdir = os.path.join(ROOT, "rga3-release_b200")
    cuda_lib = next((d for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib")
                     if os.path.exists(os.path.join(d, "libcudart.so"))), None)
    if cuda_lib is None:
        pytest.skip("libcudart.so not found")
    exe = str(tmp_path / "c_abi_demo")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", libdir, "-lb200vit", "-L", cuda_lib, "-lcudart", "-lm",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cuda_lib}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    """include/b200vit.h compiles as C99 (-pedantic: no C++ types cross the ABI) and a plain-C host program links
    against libb200vit.so and builds a plan on the host (examples/c_abi_demo.c; its device half runs in the GPU test
    tests/test_gpu_ops.py::test_c_abi_demo_runs_on_the_gpu)."""
    exe = build_c_demo(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr
    assert "b200vit version 3" in r.stdout and "launches per forward: 137" in r.stdout and "window_index holds 16384 bytes" in r.stdout
    assert "no GPU: device calls skipped" in r.stdout


def test_install_into_the_reference_unigr_wrapper():
    """SURVEY.md 8c recipe, in the build container only (needs /root/reference): the reference's own UniGRModel
    (model/qwen_2_5_vl_sam2.py:104) accepts ``install()``: the tower lands where get_video_features looks for it, keeps
    HF's state_dict keys and the attributes HF reads (dtype, spatial_merge_size).  The forward itself runs on the GPU
    against the fixture this class produced (tests/test_gpu_unigr.py)."""
    import types
    if not os.path.exists("/root/reference/model/qwen_2_5_vl_sam2.py"):
        pytest.skip("/root/reference is not available here")
    from unigr_case import unigr_config_kwargs
    saved = sys.modules.get("qwen_vl_utils")
    sys.modules["qwen_vl_utils"] = types.SimpleNamespace(process_vision_info=None)
    sys.path.insert(0, "/root/reference")
    try:
        from model.qwen_2_5_vl_sam2 import UniGRConfig, UniGRModel
        torch.manual_seed(0)
        model = UniGRModel(UniGRConfig(train_mask_decoder=True, **unigr_config_kwargs())).eval()
        keys = set(model.model.visual.state_dict().keys())
        want = {k: v.clone() for k, v in model.model.visual.state_dict().items()}
        tower = vit.install(model)                       # CPU parameters here: only the swap is exercised, no forward
        assert {k for k in tower.state_dict().keys()} == {k for k in keys if "inv_freq" not in k}
        for k, v in tower.state_dict().items():
            assert torch.equal(v, want[k]), k
        assert tower.dtype == torch.float32 and tower.spatial_merge_size == 2 and model.model.visual is tower
    finally:
        sys.path.remove("/root/reference")
        if saved is None:
            sys.modules.pop("qwen_vl_utils", None)
        else:
            sys.modules["qwen_vl_utils"] = saved
        for name in [n for n in sys.modules if n == "model" or n.startswith("model.")]:
            sys.modules.pop(name, None)
