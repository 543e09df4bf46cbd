"""Host-side logic of the drop-in classes: state_dict compatibility, BatchNorm folding and blob order
(checked by running the folded parameters through plain torch ops against the oracle), config variants."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import yolo_fastest_b200 as yf
from oracle import yolo_oracle as O
from yolo_fastest_b200.model import ARCH


def _folded_forward(blob, x, nout):
    """Forward pass from the folded blob alone (conv + bias only) — verifies fold math and canonical order."""
    pos = [0]

    def take(shape):
        n = int(np.prod(shape))
        t = torch.from_numpy(blob[pos[0]:pos[0] + n].reshape(shape).copy())
        pos[0] += n
        return t

    def conv(x, cin, cout, k, s, dw, relu):
        w = take((cout, 1 if dw else cin, k, k))
        b = take((cout,))
        y = F.conv2d(x, w, b, stride=s, padding=(k - 1) // 2, groups=cin if dw else 1)
        return F.relu(y) if relu else y

    saved = {}
    for name, kind, cin, cout, k, s, dw, relu in ARCH:
        if kind == "cbr":
            if name == "conv5_3":
                x = saved["conv5_2"]
            if name == "conv4_1_1":
                x = torch.cat((saved["conv4_2"], saved["deconv5_1"]), 1)
            x = conv(x, x.shape[1], cout, k, s, dw, relu)
            saved[name] = x
        elif kind == "res":
            y = conv(x, cin, cout, 1, 1, False, True)
            y = conv(y, cout, cout, 3, 1, True, True)
            x = conv(y, cout, cin, 1, 1, False, False) + x
        elif kind == "head":
            w = take((nout, cin, 1, 1))
            b = take((nout,))
            saved[name] = F.conv2d(x, w, b)
        else:
            w = take((cin, cout, 2, 2))
            b = take((cout,))
            saved[name] = F.relu(F.conv_transpose2d(saved["conv5_2"], w, b, stride=2))
    assert pos[0] == blob.size
    return saved["head_4"], saved["head_5"]


def test_state_dict_keys_match_shipped_checkpoints(gold):
    m = yf.YoloFastest(yf.config_params["io_params"])
    for name in ("yolo_fastest_256x320", "yolo_fastest_512x640"):
        sd = gold.sd(name)
        assert len(sd) == 508
        assert list(m.state_dict().keys()) == list(sd.keys())
        for k, v in m.state_dict().items():
            assert v.shape == sd[k].shape, k
        assert str(m.load_state_dict(sd)) == "<All keys matched successfully>"
    m80 = yf.YoloFastest({"num_cls": 80, "input_channel": 1, "num_anchors": 3})
    assert list(m80.state_dict().keys()) == list(gold.sd("stress80_416").keys())


def test_folded_blob_reproduces_oracle_forward(gold):
    for name, res in (("yolo_fastest_256x320", "256x320"), ("stress80_416", None)):
        sd = gold.sd(name)
        nc = 3 if res else 80
        m = yf.YoloFastest({"num_cls": nc, "input_channel": 1, "num_anchors": 3})
        m.load_state_dict(sd)
        m.eval()
        x = (torch.randint(0, 256, (2, 1, 64, 96), generator=torch.Generator().manual_seed(5)).float() - 128.0) / 255.0
        hl, hs = _folded_forward(m.folded_blob(), x, 3 * (5 + nc))
        rl, rs = O.forward(sd, x)
        for a, b in ((hl, rl), (hs, rs)):
            assert (a - b).abs().max().item() <= 5e-5 * b.abs().max().item() + 1e-6   # folding moves fp32 roundings
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_model_guards():
    m = yf.YoloFastest(yf.config_params["io_params"])
    assert m.training
    with pytest.raises(yf.YfError):
        m(torch.zeros(1, 1, 256, 320))                     # CPU tensor: no fallback
    m.eval()
    with pytest.raises(yf.YfError):
        m(torch.zeros(1, 1, 256, 320))
    with pytest.raises(yf.YfError):
        m.res1_1(torch.zeros(1, 4, 8, 8))                  # blocks are parameter containers
    with pytest.raises(yf.YfError):
        yf.Detect_YOLO("cpu", "unused.pth", yf.config_for("256x320"), None)
    with pytest.raises(yf.YfError):
        yf.YOLOLossV3([[10, 13]] * 3, 3, [256, 320, 1], "cpu")(torch.zeros(1, 24, 16, 20))
    with pytest.raises(yf.YfError):
        yf.non_max_suppression(torch.zeros(1, 10, 8), 3)


def test_initialize_weights_statistics():
    torch.manual_seed(0)
    m = yf.YoloFastest({"num_cls": 80, "input_channel": 1, "num_anchors": 3})
    m.initialize_weights()
    assert m.num_out == 255 and m.head_4.weight.shape == (255, 96, 1, 1)
    assert float(m.head_4.bias.abs().max()) == 0.0
    g = m.conv4_1_1[1].weight
    assert abs(float(g.mean()) - 1.0) < 0.02 and float(m.conv4_1_1[1].bias.abs().max()) == 0.0


def test_config_variants():
    a = yf.config_for("256x320")["io_params"]
    b = yf.config_for("512x640")["io_params"]
    assert a["input_shape"] == [256, 320, 1] and a["anchors"][0][0] == [10, 13] and len(a["anchors"]) == 2
    assert b["input_shape"] == [512, 640, 1] and b["anchors"][1][2] == [150, 300] and len(b["anchors"]) == 2
    assert a["conf_thre"] == 0.5 and a["nms_thre"] == 0.2 and a["class_names"][2] == "destroyer"
    assert len(yf.config_params["io_params"]["anchors"]) == 3           # untouched module-level dict
    with pytest.raises(ValueError):
        yf.config_for("1x1")


def test_bench_work_model_matches_the_survey():
    """bench.py's per-group algorithmic work (the numerators of every roofline figure) against SURVEY.md §8d: 945.8 MFLOP and
    19 763 200 B of unique fp32 activations per 640x512 image for the reference's layers (the bench counts the executed work of the
    composed heads — one 1x1 instead of two — and adds the folded weights once per launch)."""
    import bench
    for (H, W), macs_ref, act_bytes in (((512, 640), 472.9e6, 19763200), ((256, 320), 118.22e6, 4940800)):
        G = bench.group_work(H, W, 24)
        assert len(G) == 30 and [g["name"] for g in G][:3] == ["conv1_4", "res1_1", "conv2_1"]
        px = lambda d: (H // d) * (W // d)
        # the two 1x1 convs folded into the heads on the host (conv5_6: 128x128 at /32, conv4_1_5: 96x96 at /16)
        composed = px(32) * 128 * 128 + px(16) * 96 * 96
        macs = sum(g["macs"] for g in G)
        assert abs(macs + composed - macs_ref) / macs_ref < 2e-3
        # bytes: the survey's plan fuses each neck branch into one kernel; here conv5_2 / conv5_4 / conv4_1_3 are launches of their own,
        # so their outputs are written and read once more (96 + 128 channels at /32, 96 at /16), plus every folded weight once
        extra = 4 * 2 * ((96 + 128) * px(32) + 96 * px(16))
        total = sum(g["bytes"] for g in G)
        assert act_bytes <= total <= act_bytes + extra + 4 * 346356
