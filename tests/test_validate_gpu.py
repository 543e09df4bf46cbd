"""Validation driver on the GPU: the fused YF_MODE_VALIDATE kernel path against the oracle's val_decode + val_nms at
batch 16 on both resolutions, and Validation.get_mAP — forward and post-processing on the device — against the mAP frozen from the
reference's own Validation class."""
import os

import numpy as np
import pytest
import torch

import yolo_fastest_b200 as yf
from oracle import yolo_oracle as O
from yolo_fastest_b200 import _lib
from yolo_fastest_b200 import validate as V

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_fused_validate_mode_equals_oracle_at_batch_16(gold, res):
    g, cfg = gold.res[res], yf.config_for(res)
    io = cfg["io_params"]
    sd = gold.sd("yolo_fastest_" + res)
    base = g["u8"]                                                    # 20 frames at 256x320, 5 at 512x640: rolled copies make 16 distinct images
    x = torch.cat([O.preprocess_gray(np.roll(base[i % len(base)], (2 * (i // len(base)), 5 * (i // len(base))), axis=(0, 1))) for i in range(16)], 0)
    heads = O.forward(sd, x)                                          # identical inputs for both sides: the oracle's own heads
    for conf in (io["conf_thre"], 0.05):
        rows = torch.cat([O.val_decode(heads[h], io["anchors"][h], io["num_cls"], io["input_shape"]) for h in range(2)], 1)
        want = O.val_nms(rows, io["num_cls"], conf, io["nms_thre"])
        pp = yf.YOLO_post_process(conf, io["nms_thre"], io["num_anchors"], io["num_cls"], io["anchors"], io["input_shape"])
        got, counts, _ = pp._run((heads[0].cuda(), heads[1].cuda()), nms=True, mode=_lib.MODE_VALIDATE)
        assert len(got) == 16 and sum(w is not None for w in want) >= 12
        for b in range(16):
            assert len(got[b]) == (0 if want[b] is None else len(want[b])) == int(counts[b])
            if want[b] is None:
                continue
            d = got[b]
            arr = np.stack([d["x1"], d["y1"], d["x2"], d["y2"], d["conf"], d["cls_score"], d["cls"].astype(np.float64)], 1)
            assert np.array_equal(arr[:, 6], want[b][:, 6].numpy())                                   # same boxes, same order
            assert np.allclose(arr, want[b].double().numpy(), rtol=1e-5, atol=1e-5)                   # fp32 expf / sigmoid on two devices


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_get_map_equals_the_reference(gold, res):
    gm = np.load(os.path.join(gold.dir, "golden_map.npz"))
    g, cfg = gold.res[res], yf.config_for(res)
    targets = torch.from_numpy(gm["targets_" + res])
    n = targets.shape[0]
    imgs = torch.cat([O.preprocess_gray(u) for u in g["u8"][:n]], 0)
    batches = [(imgs[i:i + 5], targets[i:i + 5]) for i in range(0, n, 5)]
    lines = []

    class Log:
        def info(self, s):
            lines.append(s)
    m = yf.YoloFastest(cfg["io_params"])
    m.load_state_dict(gold.sd("yolo_fastest_" + res))
    v = V.Validation(cfg, Log(), batches, torch.device("cuda:0"))
    got = v.get_mAP(m.cuda().eval(), epoch=3)
    assert abs(got - float(gm["map_" + res])) < 1e-9 and np.allclose(v.APs, gm["aps_" + res], atol=1e-9)
    assert lines[0].endswith("epoch: 3 validation results —————") and lines[-2] == "mean AP: %.3f" % got and len(lines) == 6
