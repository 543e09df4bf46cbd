"""Phase trace of the tensor-core inverted-residual kernel (library built with -DYF_TC_TRACE, GPU box).
    YF_B200_LIB=tune/tc_trace.so python tools/tc_trace.py [res] [batch]
Prints, per chunk step of CTA 0, the cycles between the phase boundaries of worker warp 0 and of the tensor-core thread."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_fastest_b200 as yf  # noqa: E402
from yolo_fastest_b200 import _lib  # noqa: E402

res = sys.argv[1] if len(sys.argv) > 1 else "512x640"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H, W = (int(v) for v in res.split("x"))
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "yolo_fastest_%s.pth" % res), map_location="cpu")
m = yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3})
m.load_state_dict(sd)
m = m.cuda().eval()
x = ((torch.randint(0, 256, (B, 1, H, W), generator=torch.Generator().manual_seed(4)).float() - 128.0) / 255.0).cuda()
for _ in range(2):
    m(x)
torch.cuda.synchronize()
buf = (C.c_longlong * (16 * 64))()
lib = _lib.lib()
lib.yf_debug_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.yf_debug_trace(buf, 16 * 64) == 0
t = [[buf[s * 16 + e] for e in range(16)] for s in range(64)]
names = ["wait e1full", "epi1", "wait dfree/epi2", "bar1", "put_x", "dw", "bar2"]
print("step  " + " ".join("%12s" % n for n in names) + "   total |  mma: e1free->  issue1  ->dfull  issue2")
for s in range(3, 40):
    w = t[s]
    if not w[7]:
        break
    d = [w[i + 1] - w[i] for i in range(7)]
    nxt = t[s + 1][0] - w[0] if t[s + 1][0] else 0
    mm = [w[8] - w[0], w[9] - w[8], w[10] - w[0], w[11] - w[10]] if w[8] else [0, 0, w[10] - w[0], w[11] - w[10]]
    extra = " | top %d orig %d fetch %d wbar %d" % (w[12] - t[s - 1][7] if w[12] else -1, w[13] - w[12] if w[12] else -1, w[14] - w[0], w[15] - w[14])
    print("%4d  " % s + " ".join("%12d" % v for v in d) + " %7d | %8d %8d %8d %8d" % tuple([nxt] + mm) + extra)
