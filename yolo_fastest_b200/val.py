"""Drop-ins for the validation-flavour decode and NMS that validate.py:38-44 calls.

``YOLOLossV3(anchors, num_classes, input_shape, device)(head)`` mirrors the ``targets=None`` branch of
the reference loss module (src/model_training/loss/yolo_loss.py:28-36,48-68,98-141) and
``non_max_suppression(prediction, num_classes, conf_thres, nms_thres)`` mirrors
src/model_training/utils/general.py:87-143; both run as sm_100a kernels through libyf_b200.so.
The training branch of the loss (targets given) is out of scope and raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_ctx_cache = {}


def _context(device, num_cls, num_anchors, batch, ncand):
    """A post-processing-only yf_ctx with room for `ncand` candidate rows per image."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, num_cls, num_anchors)
    c = _ctx_cache.get(key)
    if c is None or c.max_batch < batch or c.ncand < ncand:
        if c is not None:
            torch.cuda.synchronize(idx)
            c.close()
        # candidate capacity of a ctx is A*(H/16*W/16 + H/32*W/32) = A*5*(H/32)*(W/32): size a square-ish dummy input to hold ncand
        cells = -(-ncand // (5 * num_anchors))
        side = int(np.ceil(np.sqrt(cells)))
        with torch.cuda.device(idx):       # post-only: a ctx allocates its forward buffers with its first weights, i.e. never here
            c = _lib.Ctx(idx, 1, num_cls, num_anchors, max(batch, 1), 32 * side, 32 * side)
        _ctx_cache[key] = c
        while len(_ctx_cache) > 4:         # a handful of (device, classes, anchors) combinations at most stay alive
            old_key = next(k for k in _ctx_cache if k != key)
            _ctx_cache.pop(old_key).close()
    return c


class YOLOLossV3(torch.nn.Module):
    def __init__(self, anchors, num_classes, input_shape, device):
        super().__init__()
        self.anchors = anchors                 # one scale: [[w, h]] * A
        self.num_anchors = len(anchors)
        self.num_classes = num_classes
        self.bbox_attrs = 5 + num_classes
        self.input_shape = input_shape
        self.device = device

    def forward(self, input, targets=None):
        if targets is not None:
            raise _lib.YfError("the training branch of YOLOLossV3 is outside the B200 detection hot path")
        if not input.is_cuda:
            raise _lib.YfError("YOLOLossV3 decode runs on CUDA tensors only (no CPU fallback)")
        x = input.contiguous().float()
        bs, ch, in_h, in_w = x.shape
        if ch != self.num_anchors * self.bbox_attrs:
            raise _lib.YfError("head has %d channels, expected %d" % (ch, self.num_anchors * self.bbox_attrs))
        n = self.num_anchors * in_h * in_w
        ctx = _context(x.device, self.num_classes, self.num_anchors, bs, n)
        out = torch.empty((bs, n, self.bbox_attrs), dtype=torch.float32, device=x.device)
        anc = (C.c_double * (2 * self.num_anchors))(*[float(v) for a in self.anchors for v in a])
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().yf_val_decode(ctx.handle, x.data_ptr(), bs, in_h, in_w, anc, self.num_anchors,
                                               self.num_classes, int(self.input_shape[0]), int(self.input_shape[1]),
                                               out.data_ptr(), C.c_void_p(stream)), ctx.handle)
        return out


def non_max_suppression(prediction, num_classes, conf_thres=0.5, nms_thres=0.4, num_anchors=3, raw=False):
    """[B, N, 5+nc] decoded rows -> list (len B) of [n, 7] fp32 tensors (x1, y1, x2, y2, obj_conf, class_conf,
    class_pred) or None, in the reference's order (classes ascending, conf descending)."""
    if not prediction.is_cuda:
        raise _lib.YfError("non_max_suppression runs on CUDA tensors only (no CPU fallback)")
    pred = prediction.contiguous().float()
    B, N, attrs = pred.shape
    if attrs != 5 + num_classes:
        raise _lib.YfError("rows have %d columns, expected %d" % (attrs, 5 + num_classes))
    ctx = _context(pred.device, num_classes, num_anchors, B, N)
    stream = torch.cuda.current_stream(pred.device).cuda_stream
    cap = min(N, 256)            # the reference's lists have no cap (general.py:87-143): repeat with the exact size when an image holds more
    while True:
        out = torch.empty((B, cap, _lib.DET_DTYPE.itemsize), dtype=torch.uint8, device=pred.device)
        counts = torch.empty((B,), dtype=torch.int32, device=pred.device)
        status = torch.empty((B,), dtype=torch.int32, device=pred.device)
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().yf_val_nms(ctx.handle, pred.data_ptr(), B, N, float(conf_thres), float(nms_thres), cap,
                                            out.data_ptr(), counts.data_ptr(), status.data_ptr(), C.c_void_p(stream)), ctx.handle)
        counts_h = counts.cpu().numpy()
        most = int(counts_h.max()) if B else 0
        if most <= cap:
            break
        cap = most
    kmax = max(most, 1)
    dets = out[:, :kmax].contiguous().cpu().numpy().view(_lib.DET_DTYPE).reshape(B, kmax)      # only the filled part travels
    result = []
    for b in range(B):
        d = dets[b, :int(counts_h[b])]
        if raw:
            result.append(d)
        elif len(d) == 0:
            result.append(None)
        else:
            rows = np.stack([d["x1"], d["y1"], d["x2"], d["y2"], d["conf"], d["cls_score"], d["cls"].astype(np.float64)], 1)
            result.append(torch.from_numpy(rows.astype(np.float32)).to(prediction.device))
    return result
