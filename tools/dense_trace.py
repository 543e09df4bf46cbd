"""Phase trace of dense_ta_kernel / dense_tc_kernel (-DYF_DENSE_TA=0) (library built with -DYF_TC_TRACE -DYF_DENSE_TRACE -DYF_TC_TRACE_CMID=1): per tile of CTA 0, cycles
between the boundaries of worker thread 0 and of the tensor-core thread."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_fastest_b200 as yf  # noqa: E402
from yolo_fastest_b200 import _lib  # noqa: E402

B, H, W = 16, 512, 640
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "yolo_fastest_512x640.pth"), map_location="cpu")
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
print("tile | x-wait  dfree0  step0  step1  step2 | tile total || mma: dfull0 issue0 dfull1 issue1 dfull2 issue2 (relative to tile start) || epilogue: ofull, tmem read, rest")
for s in range(3, 17):
    w = t[s]
    if not w[0] or not t[s + 1][0]:
        break
    print("%4d | %6d %6d %6d %8d %6d | %7d || %6d %6d %6d %6d %6d %6d || %6d %6d %6d" % (
        s, w[1] - w[0], w[6] - w[1], w[2] - w[6], w[3] - w[2], w[4] - w[3], t[s + 1][0] - w[0],
        w[8] - w[0], w[11] - w[8], w[9] - w[0], w[12] - w[9], w[10] - w[0], w[13] - w[10], w[5] - w[0], w[7] - w[5], w[14] - w[7]))
