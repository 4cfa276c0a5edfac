"""Time the STOM placement policy: device kernels (b200vit_stom_policy) vs the host policy (numpy + cv2, including
the device->host copy of the tracker outputs it needs).  python tools/bench_policy.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rga3_release_b200 as vit


def main():
    dev = "cuda"
    t, h, w, key = 16, 448, 448, 8
    res = {}
    for shape, n in (("rectangle", 1500), ("mask", 1500), ("rectangle", 8000), ("mask", 8000)):
        rng = np.random.default_rng(0)
        base = rng.uniform(120, 330, (n, 2))
        tr = np.stack([base + (i - key) * np.array([1.7, -0.9]) + rng.normal(0, 0.6, (n, 2)) for i in range(t)]).astype(np.float32)
        vi = rng.random((t, n)) < 0.9
        layer = np.zeros((h, w, 4), dtype=np.uint8)
        layer[150:300, 120:330] = (0, 255, 0, 110)
        trd, vid, layd = torch.from_numpy(tr).to(dev), torch.from_numpy(vi).to(dev), torch.from_numpy(layer).to(dev)
        for _ in range(3):
            vit.stom_frame_ops_device(trd, vid, key, shape, h, w, layd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            vit.stom_frame_ops_device(trd, vid, key, shape, h, w, layd)
        e1.record()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1) / iters
        t0 = time.perf_counter()
        for _ in range(3):
            ops = vit.stom_frame_ops(trd.cpu().numpy(), vid.cpu().numpy(), key, shape, h, w, layer)
        host_ms = (time.perf_counter() - t0) / 3 * 1e3
        res[f"{shape}_n{n}"] = {"device_ms": round(dev_ms, 4), "host_numpy_cv2_ms": round(host_ms, 3),
                                "frames": t, "points": n, "hw": [h, w]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
