"""Frame sampling / decode hand-off (rga3-release_b200/sampling.py) against vectors produced by the reference's own
utils/utils.py and utils/video_capture.py (tests/golden/make_sampling_golden.py)."""
import os
import random

import numpy as np
import pytest
import torch

import rga3_release_b200 as vit


def test_sparse_and_dense_indices_equal_the_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "sampling.npz"))
    for row in z["sparse"]:
        total, n = int(row[0]), int(row[1])
        assert vit.get_sparse_indices(total, n) == row[2:2 + n].tolist(), (total, n)
    for row in z["dense"]:
        n_mllm, n_sam = int(row[0]), int(row[1])
        assert vit.get_dense_indices(n_mllm, n_sam) == row[2:2 + n_sam].tolist(), (n_mllm, n_sam)
    assert vit.uniform_sample(10, 4) == [0, 3, 5, 8]


def test_video_capture_path_equals_the_reference(golden_dir):
    """load_frames_from_video (video_capture.py:10-60): index choice, BGR->RGB, repeat-last padding."""
    z = np.load(os.path.join(golden_dir, "sampling.npz"))
    for i in range(5):
        vlen, nf, rand, seed = (int(v) for v in z[f"vc{i}_meta"])
        idxs = vit.video_frame_indices(vlen, nf, "rand" if rand else "uniform", random.Random(seed))
        assert idxs == z[f"vc{i}_idxs"].tolist()
        decoded = []                                   # what cap.read() yields: BGR frames
        for k in range(vlen):
            f = np.zeros((6, 8, 3), dtype=np.uint8)
            f[..., 0], f[..., 1], f[..., 2] = k % 256, (k * 7) % 256, 200
            decoded.append(f)
        clip = vit.stage_clip(decoded, idxs, num_frames=nf, bgr=True)
        assert clip.shape == (nf, 6, 8, 3) and np.array_equal(clip.numpy(), z[f"vc{i}_frames"])


def test_key_frame_indices_and_staging_rules():
    idxs, pos = vit.clip_indices_with_key_frame(100, 16, 37)          # inference_videoinfer.py:77-79
    assert len(idxs) == 16 and idxs == sorted(idxs) and idxs[pos] == 37
    assert sorted(set(idxs) - {37}) == sorted(set(vit.get_sparse_indices(100, 15)) - {37})
    frames = np.arange(5 * 4 * 6 * 3, dtype=np.uint8).reshape(5, 4, 6, 3)
    buf = torch.empty((3, 4, 6, 3), dtype=torch.uint8)
    out = vit.stage_clip(frames, [4, 0, 2], out=buf)
    assert out is buf and np.array_equal(out.numpy(), frames[[4, 0, 2]])
    with pytest.raises(ValueError):
        vit.stage_clip(frames, [0, 1, 2], num_frames=2)
    with pytest.raises(ValueError):
        vit.stage_clip([np.zeros((4, 6, 3), np.uint8), np.zeros((5, 6, 3), np.uint8)], [0, 1])
