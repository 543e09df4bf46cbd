"""Per-group device time of the forward at the bench workload + a quick parity check (GPU box).
    [YF_B200_LIB=tune/lib_x.so] python tools/profile_groups.py [res] [batch] -> one JSON line"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import yolo_oracle as O  # noqa: E402
import yolo_fastest_b200 as yf  # noqa: E402

res = sys.argv[1] if len(sys.argv) > 1 else "512x640"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H, W = (int(v) for v in res.split("x"))
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "yolo_fastest_%s.pth" % res), map_location="cpu")
m = yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3})
m.load_state_dict(sd)
m = m.cuda().eval()
x = (torch.randint(0, 256, (2, 1, H, W), generator=torch.Generator().manual_seed(3)).float() - 128.0) / 255.0
ref = O.forward(sd, x)
got = m(x.cuda())
err = max(((g.cpu() - r).abs().max() / r.abs().max()).item() for g, r in zip(got, ref))
xb = (torch.randint(0, 256, (B, 1, H, W), generator=torch.Generator().manual_seed(4)).float() - 128.0) / 255.0
xb = xb.cuda()
best = None
for _ in range(3):
    p = m.profile(xb)
    best = p if best is None else [(n, min(a, b)) for (n, a), (_, b) in zip(best, p)]
print(json.dumps({"lib": os.environ.get("YF_B200_LIB", "default"), "err": err, "total_ms": sum(t for _, t in best),
                  "groups": {n: round(t, 4) for n, t in best}}))
