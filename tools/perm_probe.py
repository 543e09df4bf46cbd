"""Image independence probe (GPU box): forward of a batch and of a permutation of it, tap by tap; prints where the first difference is."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_fastest_b200 as yf  # noqa: E402

TAPS = ["conv3_4", "res3_3", "res3_4", "res3_5", "res3_6", "conv4_1"]
res = sys.argv[1] if len(sys.argv) > 1 else "256x320"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 37
H, W = (int(v) for v in res.split("x"))
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "yolo_fastest_%s.pth" % res), map_location="cpu")
m = yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3})
m.load_state_dict(sd)
m = m.cuda().eval()
x = ((torch.randint(0, 256, (B, 1, H, W), generator=torch.Generator().manual_seed(23)).float() - 128.0) / 255.0).cuda()
perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
for rep in range(3):
    m(x)
    t0 = {n: m.tap(n, B).clone() for n in TAPS}
    m(x)
    t0b = {n: m.tap(n, B).clone() for n in TAPS}
    m(x[perm])
    t1 = {n: m.tap(n, B).clone() for n in TAPS}
    for n in TAPS:
        a, b, c = t0[n], t0b[n], t1[n]
        a = a.view(B, -1, a.numel() // B // (H // 8 if n != "conv4_1" else H // 16) // (W // 8 if n != "conv4_1" else W // 16), H // 8 if n != "conv4_1" else H // 16, W // 8 if n != "conv4_1" else W // 16)[:, 0]
        b = b.view(a.shape)
        c = c.view(a.shape)
        d_rep = (a != b)
        d = (c != a[perm])
        msg = "%-8s rerun diff %d | perm diff %d of %d" % (n, int(d_rep.sum()), int(d.sum()), d.numel())
        if int(d.sum()):
            idx = d.nonzero()
            msg += " | max|d| %.3e | images %s | channels %s | rows %d..%d cols %d..%d" % (
                float((c - a[perm]).abs().max()), sorted(set(idx[:, 0].tolist()))[:8], sorted(set(idx[:, 1].tolist()))[:8],
                int(idx[:, 2].min()), int(idx[:, 2].max()), int(idx[:, 3].min()), int(idx[:, 3].max()))
        print(msg)
    print()
n = "res3_3"
a = t0[n].view(B, 16, H // 8, W // 8)
b = t0b[n].view(a.shape)
d = (a != b)
idx = d.nonzero()
import collections
print("rerun diffs res3_3: rows%8", sorted(collections.Counter((idx[:, 2] % 8).tolist()).items()))
print("cols", sorted(collections.Counter((idx[:, 3]).tolist()).items()))
print("tile rows", sorted(collections.Counter((idx[:, 2] // 8).tolist()).items()))
print("images", sorted(collections.Counter((idx[:, 0]).tolist()).items()))
print("channels", sorted(collections.Counter((idx[:, 1]).tolist()).items()))
