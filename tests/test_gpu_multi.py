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
