"""Where does the distance from exact arithmetic grow on the ill-conditioned 80-class stress network? Every tapped activation and both
heads: GPU vs float64, reference fp32 vs float64, GPU vs reference fp32 — all as max|d| / max|ref64| (GPU box).
    [YF_B200_LIB=...] python tools/stress_err.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from oracle import yolo_oracle as O  # noqa: E402
import yolo_fastest_b200 as yf  # noqa: E402
from tools.check_forward import TAPS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "stress80_416.pth"), map_location="cpu")
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
m = yf.YoloFastest({"num_cls": 80, "input_channel": 1, "num_anchors": 3})
m.load_state_dict(sd)
m = m.cuda().eval()
x = (torch.randint(0, 256, (B, 1, 416, 416), generator=torch.Generator().manual_seed(17)).float() - 128.0) / 255.0
t32, t64 = {}, {}
r32 = O.forward(sd, x, t32)
r64 = O.forward(sd64, x.double(), t64)
got = [t.cpu() for t in m(x.cuda())]
print("lib %s" % os.environ.get("YF_B200_LIB", "default"))
print("%-12s %12s %12s %12s" % ("tap", "gpu-ref64", "ref32-ref64", "gpu-ref32"))
for name in TAPS + ["head_large", "head_small"]:
    if name.startswith("head"):
        h = 0 if name == "head_large" else 1
        g, a, b = got[h].double(), r32[h].double(), r64[h]
    else:
        g, a, b = m.tap(name, B).cpu().view(t32[name].shape).double(), t32[name].double(), t64[name]
    s = b.abs().max().item()
    print("%-12s %12.3e %12.3e %12.3e" % (name, (g - b).abs().max().item() / s, (a - b).abs().max().item() / s, (g - a).abs().max().item() / s))
