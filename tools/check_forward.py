"""Per-group parity report on a GPU box: every tapped activation of yf_forward against the oracle's.

    python tools/check_forward.py [256x320|512x640|stress] [batch]
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

TAPS = ["conv1_4", "res1_1", "conv2_1", "res2_1", "res2_2", "conv3_1", "res3_1", "res3_2", "conv3_4", "res3_3", "res3_4",
        "res3_5", "res3_6", "conv4_1", "res4_1", "res4_2", "res4_3", "res4_4", "conv4_2", "conv5_1", "res5_1", "res5_2",
        "res5_3", "res5_4", "res5_5", "conv5_2", "conv5_4", "conv4_1_1", "conv4_1_3"]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "256x320"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    gold = os.path.join(ROOT, "tests", "golden")
    if which == "stress":
        sd = torch.load(os.path.join(gold, "weights", "stress80_416.pth"), map_location="cpu")
        nc, H, W = 80, 416, 416
    else:
        sd = torch.load(os.path.join(gold, "weights", "yolo_fastest_%s.pth" % which), map_location="cpu")
        nc = 3
        H, W = (256, 320) if which == "256x320" else (512, 640)
    m = yf.YoloFastest({"num_cls": nc, "input_channel": 1, "num_anchors": 3})
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = (torch.randint(0, 256, (B, 1, H, W), generator=torch.Generator().manual_seed(3)).float() - 128.0) / 255.0
    taps = {}
    rl, rs = O.forward(sd, x, taps)
    hl, hs = m(x.cuda())
    torch.cuda.synchronize()
    worst = 0.0
    print("%-10s %-22s %12s %12s" % ("tap", "shape", "max|d|/max|ref|", "max|ref|"))
    for name in TAPS:
        ref = taps[name]
        got = m.tap(name, B).cpu().view(ref.shape)
        e = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        worst = max(worst, e)
        print("%-10s %-22s %12.3e %12.4g %s" % (name, tuple(ref.shape), e, ref.abs().max().item(), "" if e < 1e-4 else "<== BAD"))
    for name, got, ref in (("head_large", hl.cpu(), rl), ("head_small", hs.cpu(), rs)):
        e = (got - ref).abs().max().item() / ref.abs().max().item()
        worst = max(worst, e)
        print("%-10s %-22s %12.3e %12.4g %s" % (name, tuple(ref.shape), e, ref.abs().max().item(), "" if e < 1e-4 else "<== BAD"))
    print("worst", worst)
    return 0 if worst < 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())
