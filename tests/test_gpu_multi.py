"""Multi-GPU sharding on real hardware (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a single-GPU box).  The work is tools/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clip_and_slice_sharding_match_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_tower_on_a_second_gpu_in_the_same_process():
    """Per-device state (function attributes, the normalisation LUT symbol, SM count, L2 carve-out) is keyed by device:
    a tower on cuda:1 next to one on cuda:0 gives the same embeddings, while cuda:0 stays the current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import numpy as np
    sys.path.insert(0, ROOT)
    import rga3_release_b200 as vit
    from rga3_release_b200 import _lib
    from oracle import hf_ref, overlay_ref, tower_ref
    cfg = tower_ref.TowerCfg(**hf_ref.CFG_SMALL)
    sd = hf_ref.make_state_dict(cfg, 3)
    frames = hf_ref.synthetic_frames(4, 112, 84, clip_id=2)
    layer = overlay_ref.box_layer_ref(112, 84, (10, 12, 70, 90), 3, (255, 0, 0, 200))
    ops = [vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i, sy=-i) for i in range(4)]
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        t = vit.B200VisionTower(dict(hf_ref.CFG_SMALL), device=dev, return_dict=False)
        t.load_state_dict(sd)
        o = t.forward_frames(frames.to(dev), vit.OverlaySpec.from_rgba(layer, ops, device=dev))
        assert o.device == torch.device(dev) and torch.cuda.current_device() == 0
        outs.append(o.cpu())
        with pytest.raises(ValueError):                     # no silent cross-device work
            t.forward_frames(frames.to("cuda:1" if dev == "cuda:0" else "cuda:0"))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    t1 = vit.B200VisionTower(dict(hf_ref.CFG_SMALL), device="cuda:0", return_dict=False)
    t1.load_state_dict(sd)
    t1.to("cuda:1")                                         # .to(device) moves plans, workspaces and packed weights
    assert torch.equal(t1.forward_frames(frames.to("cuda:1"), vit.OverlaySpec.from_rgba(layer, ops, device="cuda:1")).cpu(), outs[0])
