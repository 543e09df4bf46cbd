"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference

Imports the reference's own modules from /root/reference (never copies its sources), runs its
Detect_YOLO pre-process -> YoloFastest.forward -> YOLO_post_process pipeline (src/detect.py) and
its validation-flavour decode/NMS (yolo_loss.py, general.py) on the 20 shipped test images and on
seeded synthetic inputs, asserts that oracle/yolo_oracle.py reproduces every output bit for bit,
and writes the vectors to tests/golden/.  The GPU box has no /root/reference: tests there read
only the files written here.

Fixtures written (data, not code): the 20 test JPEGs, the two shipped checkpoints, a calibrated
random-init 80-class checkpoint for the NMS stress config, and golden_*.npz.
"""
import copy
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(REF, "src", "model_training"))
sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))  # train.py:11, absent here

import cv2  # noqa: E402
import detect as ref_detect  # noqa: E402  (src/detect.py)
from model_training._config import config_params as ref_config  # noqa: E402
from model.yolo_fastest import YoloFastest as RefYoloFastest  # noqa: E402
from loss.yolo_loss import YOLOLossV3 as RefYOLOLossV3  # noqa: E402
from utils.general import non_max_suppression as ref_val_nms  # noqa: E402

from oracle import yolo_oracle as O  # noqa: E402

ref_detect.device = torch.device("cpu")  # module global read by decode_box (detect.py:44)
torch.set_grad_enabled(False)

CKPT = {
    "256x320": os.path.join(REF, "models/pytorch/256x320/YOLO-Fastest_epoch_28.pth"),
    "512x640": os.path.join(REF, "models/pytorch/512x640/YOLO-Fastest_epoch_27.pth"),
}
COCO_ANCHORS = [[[12, 18], [37, 49], [52, 132]], [[115, 73], [119, 199], [242, 238]]]  # yolo_fastest.py:403-406


def cfg_for(res):
    cfg = copy.deepcopy(ref_config)
    io = cfg["io_params"]
    if res == "256x320":
        io["input_shape"] = [256, 320, 1]
        io["anchors"] = io["anchors"][0:2]      # "256x320 uses the first two" (_config.py:5-9)
    else:
        io["input_shape"] = [512, 640, 1]
        io["anchors"] = io["anchors"][1:3]      # "512x640 uses the last two"
    return cfg


class _NullLogger:
    def info(self, *a, **k):
        pass


def rows_to_array(rows, with_src=False):
    n = 8 if with_src else 7
    a = np.zeros((len(rows), n), dtype=np.float64)
    for i, r in enumerate(rows):
        a[i, :len(r)] = [float(v) for v in r]
    return a


def ref_postprocess(pp, pred, num_cls):
    """detect.py:155-169 driven through the reference's own YOLO_post_process object."""
    boxes = pp.decode_box(pred)
    decoded = copy.deepcopy(boxes)
    per_cls = [[] for _ in range(num_cls)]
    for b in boxes:
        per_cls[b[-1]].append(b)
    out = []
    for c in range(num_cls):
        if not per_cls[c]:
            continue
        per_cls[c].sort(key=lambda item: item[4], reverse=True)
        out.extend(pp.non_maxium_supression(per_cls[c]))
    return decoded, out


def same_rows(a, b):
    if len(a) != len(b):
        return False
    for ra, rb in zip(a, b):
        if [float(v) for v in ra[:7]] != [float(v) for v in rb[:7]]:
            return False
    return True


def do_resolution(res, names):
    cfg = cfg_for(res)
    io = cfg["io_params"]
    det = ref_detect.Detect_YOLO(torch.device("cpu"), CKPT[res], cfg, _NullLogger())
    sd = torch.load(CKPT[res], map_location="cpu")
    H, W = io["input_shape"][0:2]
    u8s, hl_all, hs_all = [], [], []
    dec_all, kept_all, kept_adj_all, val_all = [], [], [], []
    flags = []
    losses = [RefYOLOLossV3(io["anchors"][i], io["num_cls"], io["input_shape"], torch.device("cpu")) for i in range(2)]
    for name in names:
        img, _ = det._Detect_YOLO__pre_process(os.path.join(REF, "test_data", name))
        u8 = (img * 255.0 + 128.0).round().to(torch.uint8)[0, 0].numpy()
        assert torch.equal(O.preprocess_gray(u8), img), "oracle preprocess != reference"
        pred = det.model(img)
        o_pred = O.forward(sd, img)
        assert torch.equal(pred[0], o_pred[0]) and torch.equal(pred[1], o_pred[1]), "oracle forward != reference"
        decoded, kept = ref_postprocess(det.post_process, pred, io["num_cls"])
        o_dec = O.decode_box(pred, io["anchors"], io["input_shape"], io["conf_thre"], io["num_anchors"], io["num_cls"])
        o_kept = O.detect_postprocess(pred, io["anchors"], io["input_shape"], io["conf_thre"], io["nms_thre"],
                                      io["num_anchors"], io["num_cls"])
        assert same_rows(decoded, o_dec), "oracle decode != reference"
        assert same_rows(kept, o_kept), "oracle nms != reference"
        kept_adj = copy.deepcopy(kept)
        o_adj = copy.deepcopy(o_kept)
        if io["input_shape"][0:2] != io["origin_img_shape"][0:2]:
            det._Detect_YOLO__adjust_coord(kept_adj)
            O.adjust_coord(o_adj, io["input_shape"], io["origin_img_shape"])
        assert same_rows(kept_adj, o_adj)
        # validation flavour (validate.py:38-44)
        v = torch.cat([losses[i](pred[i]) for i in range(2)], 1)
        o_v = torch.cat([O.val_decode(pred[i], io["anchors"][i], io["num_cls"], io["input_shape"]) for i in range(2)], 1)
        assert torch.equal(v, o_v), "oracle val_decode != reference"
        r_out = ref_val_nms(v.clone(), io["num_cls"], conf_thres=io["conf_thre"], nms_thres=io["nms_thre"])[0]
        o_out = O.val_nms(o_v, io["num_cls"], io["conf_thre"], io["nms_thre"])[0]
        assert (r_out is None) == (o_out is None)
        if r_out is not None:
            assert torch.equal(r_out, o_out), "oracle val_nms != reference"
        u8s.append(u8)
        hl_all.append(pred[0][0].numpy())
        hs_all.append(pred[1][0].numpy())
        dec_all.append(rows_to_array(o_dec))
        kept_all.append(rows_to_array(kept))
        kept_adj_all.append(rows_to_array(kept_adj))
        val_all.append(np.zeros((0, 7), np.float32) if r_out is None else r_out.numpy())
        flags.append(len(kept) > 0)
    # seeded synthetic images (uniform u8 pixels), checked oracle == reference, stored as golden heads
    g = torch.Generator().manual_seed(1234)
    syn_u8 = torch.randint(0, 256, (2, H, W), generator=g, dtype=torch.uint8).numpy()
    syn_x = torch.cat([O.preprocess_gray(s) for s in syn_u8], 0)
    r = det.model(syn_x)
    o = O.forward(sd, syn_x)
    assert torch.equal(r[0], o[0]) and torch.equal(r[1], o[1])
    keep_heads = len(names) if res == "256x320" else 5   # keep the 512x640 fixture small
    out = {
        "names": np.array(names),
        "u8": np.stack(u8s[:keep_heads]),
        "head_large": np.stack(hl_all[:keep_heads]),
        "head_small": np.stack(hs_all[:keep_heads]),
        "syn_seed": np.array(1234),
        "syn_head_large": r[0].numpy(),
        "syn_head_small": r[1].numpy(),
        "has_targets": np.array(flags),
    }
    for i in range(len(names)):
        out["decoded_%02d" % i] = dec_all[i]
        out["kept_%02d" % i] = kept_all[i]
        out["kept_adj_%02d" % i] = kept_adj_all[i]
        out["val_%02d" % i] = val_all[i]
    np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % res), **out)
    print(res, "detect flags:", "".join("1" if f else "0" for f in flags))
    return flags


def make_stress():
    """80-class, 416x416, conf 0.001 dense-box NMS stress config."""
    torch.manual_seed(0)
    io = {"num_cls": 80, "input_channel": 1, "num_anchors": 3}
    m = RefYoloFastest(io)
    m.initialize_weights()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.momentum = None            # cumulative average => running stats = batch stats
    m.train()
    g = torch.Generator().manual_seed(7)
    for _ in range(2):
        x = (torch.randint(0, 256, (8, 1, 416, 416), generator=g).float() - 128.0) / 255.0
        m(x)
    m.eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.save(sd, os.path.join(HERE, "weights", "stress80_416.pth"))
    u8 = torch.randint(0, 256, (2, 416, 416), generator=torch.Generator().manual_seed(99), dtype=torch.uint8).numpy()
    x = torch.cat([O.preprocess_gray(s) for s in u8], 0)
    pred = m(x)
    o = O.forward(sd, x)
    assert torch.equal(pred[0], o[0]) and torch.equal(pred[1], o[1])
    pp = ref_detect.YOLO_post_process(0.001, 0.2, 3, 80, COCO_ANCHORS, [416, 416, 1])
    out = {"u8_seed": np.array(99), "head_large": pred[0].numpy(), "head_small": pred[1].numpy()}
    for b in range(2):
        pb = (pred[0][b:b + 1], pred[1][b:b + 1])
        try:
            decoded, kept = ref_postprocess(pp, pb, 80)
            ok = True
        except ZeroDivisionError:
            ok = False          # reference raises on 0/0; oracle defines NaN -> keep
        o_kept = O.detect_postprocess(pb, COCO_ANCHORS, [416, 416, 1], 0.001, 0.2, 3, 80)
        if ok:
            assert same_rows(kept, o_kept)
        out["kept_%02d" % b] = rows_to_array(o_kept)
        out["ref_raised_%02d" % b] = np.array(not ok)
        print("stress image", b, "survivors", len(O.decode_box(pb, COCO_ANCHORS, [416, 416, 1], 0.001, 3, 80)),
              "kept", len(o_kept), "reference raised" if not ok else "")
    np.savez_compressed(os.path.join(HERE, "golden_stress80_416.npz"), **out)


def main():
    os.makedirs(os.path.join(HERE, "images"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "weights"), exist_ok=True)
    names = sorted(os.listdir(os.path.join(REF, "test_data")))
    for n in names:
        shutil.copyfile(os.path.join(REF, "test_data", n), os.path.join(HERE, "images", n))
    for res, p in CKPT.items():
        shutil.copyfile(p, os.path.join(HERE, "weights", "yolo_fastest_%s.pth" % res))
    f1 = do_resolution("256x320", names)
    f2 = do_resolution("512x640", names)
    # published pins: test_result/*/笔记本cpu(python)_test_result/cpu-test.log
    assert all(f1), "256x320: every image has detections in the shipped log"
    assert [n for n, f in zip(names, f2) if not f] == ["noCloud_2m_4359.jpg"], "512x640: only noCloud_2m_4359 has none"
    make_stress()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
