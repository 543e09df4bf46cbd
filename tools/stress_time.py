"""Timing of BASELINE configuration 5 (80-class head, 416x416, conf 0.001: every one of the 2 535 candidates per image survives the
confidence filter): forward, and the head kernel ALONE (yf_decode = phase 1 only; yf_postprocess = decode + sort + NMS) through the C
ABI on device buffers, CUDA events, no host copies; then the public call (`postprocess_batch`, with its result copy)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_fastest_b200 as yf  # noqa: E402
from yolo_fastest_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "stress80_416.pth"), map_location="cpu")
m = yf.YoloFastest({"num_cls": 80, "input_channel": 1, "num_anchors": 3})
m.load_state_dict(sd)
m = m.cuda().eval()
x = ((torch.randint(0, 256, (B, 1, 416, 416), generator=torch.Generator().manual_seed(3)).float() - 128.0) / 255.0).cuda()
pp = yf.YOLO_post_process(0.001, 0.2, 3, 80, yf.COCO_ANCHORS, [416, 416, 1])
hl, hs = m(x)
ctx = pp._context(hl.device, B)
cap = 2535
out = torch.empty((B, cap, 56), dtype=torch.uint8, device="cuda")
cnt = torch.empty(B, dtype=torch.int32, device="cuda")
st = torch.empty(B, dtype=torch.int32, device="cuda")
p = pp._params(_lib.MODE_DETECT, cap)
lib = _lib.lib()


def run(fn):
    _lib.check(fn(ctx.handle, hl.data_ptr(), hs.data_ptr(), B, 26, 26, 13, 13, C.byref(p), out.data_ptr(), cnt.data_ptr(), st.data_ptr(), None), ctx.handle)


def timed(f, reps=10):
    for _ in range(3):
        f()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        f()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


t_fwd = timed(lambda: m(x))
t_dec = timed(lambda: run(lib.yf_decode))
t_all = timed(lambda: run(lib.yf_postprocess))
kept = float(cnt.float().mean())
t_api = timed(lambda: pp.postprocess_batch((hl, hs), raw=True), reps=3)
print("batch %d, 2535 candidates per image, 80 classes: forward %.3f ms | head kernel: decode only %.3f ms, decode + sort + NMS %.3f ms "
      "(%.1f boxes kept per image) | postprocess_batch incl. result copy %.3f ms" % (B, t_fwd, t_dec, t_all, kept, t_api))
