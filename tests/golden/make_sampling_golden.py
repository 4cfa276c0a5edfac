"""Golden frame-sampling vectors from the REFERENCE's own functions: /root/reference/utils/utils.py (uniform_sample,
get_sparse_indices, get_dense_indices) and /root/reference/utils/video_capture.py (load_frames_from_video, driven by a
fake cv2.VideoCapture that serves synthetic frames), imported unmodified (matplotlib stubbed).
  python tests/golden/make_sampling_golden.py   -> tests/golden/sampling.npz"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
sys.path.insert(0, "/root/reference")
import cv2  # noqa: E402
from utils import utils as ref_utils  # noqa: E402
from utils import video_capture as ref_vc  # noqa: E402


class FakeCapture:
    """cv2.VideoCapture stand-in: frame i is a 6x8 BGR image whose pixels encode i."""
    def __init__(self, path):
        self.n = int(path)
        self.pos = 0

    def isOpened(self):
        return True

    def get(self, prop):
        return float(self.n)

    def set(self, prop, value):
        self.pos = int(value)

    def read(self):
        if self.pos >= self.n:
            return False, None
        f = np.zeros((6, 8, 3), dtype=np.uint8)
        f[..., 0], f[..., 1], f[..., 2] = self.pos % 256, (self.pos * 7) % 256, 200
        self.pos += 1
        return True, f

    def release(self):
        pass


def main():
    out = {}
    sparse, dense = [], []
    for total in list(range(1, 40)) + [64, 100, 257, 1000]:
        for n in (1, 2, 3, 4, 7, 8, 15, 16, 31, 32):
            sparse.append([total, n] + ref_utils.get_sparse_indices(total, n) + [-1] * (40 - n))
    for n_mllm in range(2, 40):
        for n_sam in (1, 2, 4, 8):
            dense.append([n_mllm, n_sam] + ref_utils.get_dense_indices(n_mllm, n_sam) + [-1] * (8 - n_sam))
    out["sparse"] = np.array(sparse, dtype=np.int64)
    out["dense"] = np.array(dense, dtype=np.int64)
    ref_vc.cv2.VideoCapture = FakeCapture
    rows = []
    for vlen, nf, mode, seed in [(100, 8, "uniform", 0), (5, 8, "uniform", 0), (37, 16, "rand", 3), (16, 16, "uniform", 0),
                                 (300, 32, "rand", 11)]:
        random.seed(seed)
        frames, idxs = ref_vc.VideoCapture.load_frames_from_video(str(vlen), nf, sample=mode)
        assert len(frames) == nf
        rows.append(dict(vlen=vlen, nf=nf, mode=mode, seed=seed, idxs=idxs, frames=np.stack(frames)))
    for i, r in enumerate(rows):
        out[f"vc{i}_meta"] = np.array([r["vlen"], r["nf"], 1 if r["mode"] == "rand" else 0, r["seed"]])
        out[f"vc{i}_idxs"] = np.array(r["idxs"], dtype=np.int64)
        out[f"vc{i}_frames"] = r["frames"]
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), **out)
    print("wrote", len(sparse), "sparse,", len(dense), "dense,", len(rows), "capture cases")


if __name__ == "__main__":
    main()
