"""Where a tapped activation differs from the oracle's: error per channel and per pixel (GPU box).

    python tools/diag_tap.py conv5_1 [256x320|512x640] [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from oracle import yolo_oracle as O  # noqa: E402
import yolo_fastest_b200 as yf  # noqa: E402


def main():
    name = sys.argv[1]
    which = sys.argv[2] if len(sys.argv) > 2 else "256x320"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "yolo_fastest_%s.pth" % which), map_location="cpu")
    H, W = (256, 320) if which == "256x320" else (512, 640)
    m = yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3})
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = (torch.randint(0, 256, (B, 1, H, W), generator=torch.Generator().manual_seed(3)).float() - 128.0) / 255.0
    taps = {}
    O.forward(sd, x, taps)
    m(x.cuda())
    torch.cuda.synchronize()
    ref = taps[name]
    got = m.tap(name, B).cpu().view(ref.shape)
    d = (got - ref).abs()
    torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
    print("max|ref|", ref.abs().max().item(), "nan", torch.isnan(got).sum().item())
    print("per channel max|d|:", d.amax(dim=(0, 2, 3)))
    print("per pixel max|d| (image 0):")
    print(d[0].amax(dim=0))
    print("got[0, 0]:")
    print(got[0, 0])
    print("ref[0, 0]:")
    print(ref[0, 0])


if __name__ == "__main__":
    main()
