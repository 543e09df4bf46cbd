"""Phase trace of the channel-lane tensor-core kernel (library built with -DYF_TC_TRACE -DYF_TCT_TRACE -DYF_TC_TRACE_CMID=1, GPU box).
    YF_B200_LIB=tune/tct_trace.so python tools/tct_trace.py [res] [batch]
Cycles per tile of CTA 0: worker warp 0, the tensor-core thread and staging warp 0, all relative to the worker's tile start."""
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
print("tile | worker: efull-wait  compute  dfree-wait  stores  fence+arrive | total || mma (rel. to worker tile start): expand(t+1) at, issue; project(t) at, issue || staging of tile t (rel.): start, rawfull, xfree, put done")
for s in range(3, 24):
    w = t[s]
    if not w[0] or not t[s + 1][0]:
        break
    o = w[0]
    print("%4d | %10d %8d %10d %8d %8d | %6d || %6d %6d %6d %6d || %6d %6d %6d %6d" % (
        s, w[1] - w[0], w[2] - w[1], w[3] - w[2], w[5] - w[3], w[4] - w[5], t[s + 1][0] - w[0],
        w[6] - o, w[7] - w[6], w[8] - o, w[9] - w[8], w[10] - o, w[11] - o, w[12] - o, w[13] - o))
