"""Timing of BASELINE configuration 5 (80-class head, 416x416, conf 0.001: every one of the 2 535 candidates per image survives the
confidence filter): forward + head kernel, and the head kernel alone, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_fastest_b200 as yf  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = torch.load(os.path.join(ROOT, "tests", "golden", "weights", "stress80_416.pth"), map_location="cpu")
m = yf.YoloFastest({"num_cls": 80, "input_channel": 1, "num_anchors": 3})
m.load_state_dict(sd)
m = m.cuda().eval()
x = ((torch.randint(0, 256, (B, 1, 416, 416), generator=torch.Generator().manual_seed(3)).float() - 128.0) / 255.0).cuda()
pp = yf.YOLO_post_process(0.001, 0.2, 3, 80, yf.COCO_ANCHORS, [416, 416, 1])
pred = m(x)
for max_det in (2535, 300):
    for _ in range(3):
        out = pp.postprocess_batch(pred, max_det=max_det, raw=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(5):
        pred = m(x)
    ev[1].record()
    for _ in range(5):
        pp._run(pred, True, max_det=max_det)
    ev[2].record()
    torch.cuda.synchronize()
    kept = sum(len(o) for o in out) / B
    print("batch %d max_det %d: forward %.3f ms, decode + 80-class NMS %.3f ms (%.1f boxes kept per image of 2535 candidates)"
          % (B, max_det, ev[0].elapsed_time(ev[1]) / 5, ev[1].elapsed_time(ev[2]) / 5, kept))
