"""Time the resize step: python tools/bench_resize.py  (16 frames 720p/1080p -> smart_resize target)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rga3_release_b200 as vit


def main():
    res = {}
    for h, w, mp in ((720, 1280, 384 * 28 * 28), (1080, 1920, 768 * 28 * 28), (480, 854, 256 * 28 * 28)):
        t = 16
        fr = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, device="cuda")
        oh, ow = vit.smart_resize(h, w, 28, 4 * 28 * 28, mp)
        out = torch.empty((t, oh, ow, 3), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            vit.resize_frames(fr, oh, ow, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            vit.resize_frames(fr, oh, ow, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        byts = t * 3 * (h * w + 2 * h * ow + oh * ow)       # in + intermediate written and read + out
        res[f"{h}x{w}->{oh}x{ow}"] = {"ms": round(ms, 4), "algorithmic_GBps": round(byts / ms / 1e6, 1), "frames": t}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
