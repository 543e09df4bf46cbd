"""How far is the GPU result from the reference's fp32 result, compared with how far the reference's
own fp32 arithmetic is from exact (fp64) arithmetic on the same weights and inputs?

    python tools/noise_floor.py          (GPU box)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from oracle import yolo_oracle as O  # noqa: E402
import yolo_fastest_b200 as yf  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def stats(name, a, b):
    d = (a.double() - b.double()).abs()
    scale = b.double().abs().max().item()
    viol = (d > 1e-4 + 1e-4 * b.double().abs()).sum().item()
    i = d.argmax().item()
    print("  %-34s max|d| %.3e  /max|ref| %.3e  rms %.3e  allclose(1e-4,1e-4) violations %d/%d  worst ref value %.4g"
          % (name, d.max().item(), d.max().item() / scale, d.pow(2).mean().sqrt().item(), viol, d.numel(), b.flatten()[i].item()))


def run(tag, ckpt, nc, x):
    sd = torch.load(os.path.join(GOLD, "weights", ckpt + ".pth"), map_location="cpu")
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    m = yf.YoloFastest({"num_cls": nc, "input_channel": 1, "num_anchors": 3})
    m.load_state_dict(sd)
    m = m.cuda().eval()
    r32 = O.forward(sd, x)
    r64 = O.forward(sd64, x.double())
    torch.set_num_threads(1)
    r32_1t = O.forward(sd, x)
    torch.set_num_threads(os.cpu_count())
    got = [t.cpu() for t in m(x.cuda())]
    print(tag)
    for h, hn in enumerate(("head_large", "head_small")):
        stats(hn + " gpu vs ref32", got[h], r32[h])
        stats(hn + " gpu vs ref64", got[h], r64[h])
        stats(hn + " ref32 vs ref64", r32[h], r64[h])
        stats(hn + " ref32(1 thread) vs ref32", r32_1t[h], r32[h])


def main():
    for res in ("256x320", "512x640"):
        g = np.load(os.path.join(GOLD, "golden_%s.npz" % res))
        x = torch.cat([O.preprocess_gray(u) for u in g["u8"][:5]], 0)
        run("shipped images " + res, "yolo_fastest_" + res, 3, x)
    x = (torch.randint(0, 256, (2, 1, 416, 416), generator=torch.Generator().manual_seed(17)).float() - 128.0) / 255.0
    run("stress80 416x416 random", "stress80_416", 80, x)


if __name__ == "__main__":
    main()
