"""Pin oracle/map_oracle.py against the UNMODIFIED reference's Validation class (build container only; needs /root/reference).

    python tests/golden/make_golden_map.py

Ground truth does not ship with the reference, so it is synthesised from the reference's own validation-flavour detections on the
20 shipped images (golden_<res>.npz heads): most boxes are jittered by a few pixels (still IoU > 0.5), some are moved away (false
positive + missed target), some get a wrong class, some extra targets are added.  The reference's `Validation.get_mAP` is driven
with its own YOLOLossV3 + non_max_suppression on those heads (the "model" returns the frozen heads); the oracle must give the same
per-class APs and mAP.  Writes tests/golden/golden_map.npz (targets + expected numbers)."""
import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src", "model_training"))
sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))

from _config import config_params as ref_config  # noqa: E402
from loss.yolo_loss import YOLOLossV3 as RefLoss  # noqa: E402
from validate import Validation as RefValidation  # noqa: E402
from oracle import map_oracle as M  # noqa: E402
from oracle import yolo_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


class Log:
    def __init__(self):
        self.lines = []

    def info(self, s):
        self.lines.append(s)

    error = info


out = {}
for res in ("256x320", "512x640"):
    g = np.load(os.path.join(HERE, "golden_%s.npz" % res))
    cfg = copy.deepcopy(ref_config)
    io = cfg["io_params"]
    io["input_shape"], io["anchors"] = ([256, 320, 1], io["anchors"][0:2]) if res == "256x320" else ([512, 640, 1], io["anchors"][1:3])
    heads = (torch.from_numpy(g["head_large"]), torch.from_numpy(g["head_small"]))
    n = heads[0].shape[0]
    bs = 5
    # synthetic ground truth from the reference's own detections
    rows = torch.cat([O.val_decode(heads[h], io["anchors"][h], io["num_cls"], io["input_shape"]) for h in range(2)], 1)
    dets = O.val_nms(rows, io["num_cls"], io["conf_thre"], io["nms_thre"])
    rng = np.random.default_rng(7 if res == "256x320" else 8)
    H, W = io["input_shape"][0:2]
    targets = np.zeros((n, 64, 6), dtype=np.float32)
    for b in range(n):
        k = 0
        for d in ([] if dets[b] is None else dets[b].numpy()):
            x1, y1, x2, y2, cls = d[0], d[1], d[2], d[3], int(d[6])
            u = rng.random()
            if u < 0.6:
                dx, dy = rng.integers(-2, 3, 2)
            elif u < 0.75:
                dx, dy = 40, 30                                   # moved away: the prediction is a false positive, the target is missed
            elif u < 0.85:
                dx, dy, cls = 0, 0, (cls + 1) % io["num_cls"]     # wrong class
            else:
                continue                                          # no target at all for this prediction
            bx = [(x1 + x2) / 2 + dx, (y1 + y2) / 2 + dy, x2 - x1, y2 - y1]
            targets[b, k] = [bx[0] / W, bx[1] / H, bx[2] / W, bx[3] / H, cls, 255]
            k += 1
        if rng.random() < 0.3:                                    # an extra target nothing predicts
            targets[b, k] = [0.1, 0.1, 0.05, 0.05, int(rng.integers(0, io["num_cls"])), 255]
    targets = torch.from_numpy(targets)
    # the reference class, driven without its DataLoader
    log = Log()
    v = RefValidation.__new__(RefValidation)
    v.logger, v.device, v.bs = log, torch.device("cpu"), bs
    v.model_loss = [RefLoss(io["anchors"][h], io["num_cls"], io["input_shape"], torch.device("cpu")) for h in range(2)]
    v.input_shape, v.num_cls, v.cls_name = io["input_shape"], io["num_cls"], io["class_names"]
    v.IOU_threshold, v.conf_thres, v.nms_thres = cfg["train_params"]["IOU_val_thre"], io["conf_thre"], io["nms_thre"]
    v.target_num = torch.zeros((v.num_cls))
    v.match_list = [[] for _ in range(v.num_cls)]
    v.dataloader = [(torch.arange(i, i + bs), targets[i:i + bs].clone()) for i in range(0, n, bs)]

    class FakeModel:
        def eval(self):
            pass

        def __call__(self, imgs):
            idx = imgs.long()
            return heads[0][idx], heads[1][idx]

    ref_map = float(v.get_mAP(FakeModel(), 0))
    ref_aps = [float(l.split("AP = ")[1]) for l in log.lines if "AP = " in l]
    batches = []
    for i in range(0, n, bs):
        r = torch.cat([O.val_decode(heads[h][i:i + bs], io["anchors"][h], io["num_cls"], io["input_shape"]) for h in range(2)], 1)
        batches.append((O.val_nms(r, io["num_cls"], io["conf_thre"], io["nms_thre"]), targets[i:i + bs]))
    o_map, o_aps, o_tn = M.get_map(batches, io["num_cls"], io["input_shape"], cfg["train_params"]["IOU_val_thre"])
    print(res, "reference mAP %.6f" % ref_map, ref_aps, "| oracle %.6f" % o_map, ["%.3f" % a for a in o_aps], o_tn)
    # the reference accumulates recall through float32 tensors (its target_num is one): agreement to float32 precision
    assert abs(ref_map - o_map) < 1e-6 and all(abs(a - round(b, 3)) < 5.1e-4 for a, b in zip(ref_aps, o_aps))
    assert [float(x) for x in v.target_num] == o_tn
    out["targets_" + res], out["map_" + res], out["aps_" + res] = targets.numpy(), np.float64(o_map), np.array(o_aps)
np.savez_compressed(os.path.join(HERE, "golden_map.npz"), **out)
print("oracle mAP == reference Validation.get_mAP; wrote golden_map.npz")
