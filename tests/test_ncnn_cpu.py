"""ncnn weight source: the reference's shipped models/ncnn/256x320 files (copied to tests/golden/ncnn) parse, match the
architecture and carry the same folded parameters as the shipped .pth (ncnnoptimize folded the BatchNorms in fp32)."""
import os

import numpy as np
import pytest
import torch

import yolo_fastest_b200 as yf
from yolo_fastest_b200 import ncnn_loader

HERE = os.path.dirname(os.path.abspath(__file__))
PARAM = os.path.join(HERE, "golden", "ncnn", "YOLO-Fastest_epoch_28-opt.param")
BIN = os.path.join(HERE, "golden", "ncnn", "YOLO-Fastest_epoch_28-opt.bin")


def _model():
    return yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3})


def test_layer_table():
    layers = ncnn_loader.read_ncnn(PARAM, BIN)
    assert len(layers) == 86                                            # 86 convolutions
    assert sum(l["type"] == "Deconvolution" for l in layers) == 1 and sum(l["type"] == "ConvolutionDepthWise" for l in layers) == 27
    assert layers[0]["num_output"] == 8 and layers[0]["kernel"] == 3 and layers[0]["stride"] == 2
    assert layers[-1]["num_output"] == 24 and all(l["bias"] is not None for l in layers)


def test_folded_parameters_equal_the_checkpoint():
    a = _model().load_ncnn(PARAM, BIN).folded_blob()
    m = _model()
    m.load_state_dict(torch.load(os.path.join(HERE, "golden", "weights", "yolo_fastest_256x320.pth"), map_location="cpu"))
    b = m.folded_blob()
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= 2e-6 * float(np.abs(b).max())   # two fp32 foldings of the same BatchNorms
    # and the ncnn arrays pass through folded_blob unchanged (identity BatchNorms)
    raw = []
    for l in ncnn_loader.read_ncnn(PARAM, BIN):
        w = l["weight"]
        if l["type"] == "Deconvolution":
            co, k = l["num_output"], l["kernel"]
            w = w.reshape(co, -1, k, k).transpose(1, 0, 2, 3).ravel()
        raw += [w, l["bias"]]
    assert np.array_equal(a, np.concatenate(raw))
    # a checkpoint loaded afterwards folds with the reference's eps again
    m2 = _model().load_ncnn(PARAM, BIN)
    m2.load_state_dict(torch.load(os.path.join(HERE, "golden", "weights", "yolo_fastest_256x320.pth"), map_location="cpu"))
    assert np.array_equal(m2.folded_blob(), b)


def test_rejects_foreign_files(tmp_path):
    bad = tmp_path / "x.param"
    bad.write_text("12345\n1 1\n")
    with pytest.raises(yf.YfError):
        ncnn_loader.read_ncnn(str(bad), BIN)
    # a truncated .param (fewer layers than the .bin holds)
    lines = open(PARAM).read().split("\n")
    short = tmp_path / "short.param"
    short.write_text("\n".join(lines[:40]))
    with pytest.raises(yf.YfError):
        ncnn_loader.read_ncnn(str(short), BIN)
