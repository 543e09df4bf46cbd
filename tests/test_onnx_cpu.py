"""ONNX weight source: the reference's shipped models/onnx/256x320 export (copied to tests/golden/onnx) parses with the
package's own protobuf reader, matches the architecture and carries exactly the tensors of the shipped .pth."""
import os

import numpy as np
import pytest
import torch

import yolo_fastest_b200 as yf
from yolo_fastest_b200 import onnx_loader

HERE = os.path.dirname(os.path.abspath(__file__))
ONNX = os.path.join(HERE, "golden", "onnx", "YOLO-Fastest_epoch_28.onnx")
PTH = os.path.join(HERE, "golden", "weights", "yolo_fastest_256x320.pth")


def _model():
    return yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3})


def test_node_table():
    layers = onnx_loader.read_onnx(ONNX)
    assert len(layers) == 86                                            # 86 convolutions
    assert sum(l["type"] == "ConvTranspose" for l in layers) == 1 and sum(l["bn"] is not None for l in layers) == 84
    assert sum(l["group"] > 1 for l in layers) == 27                    # depthwise convolutions
    assert layers[0]["num_output"] == 8 and layers[0]["kernel"] == 3 and layers[0]["stride"] == 2
    heads = [l for l in layers if l["bn"] is None]
    assert [h["num_output"] for h in heads] == [24, 24] and all(h["bias"] is not None for h in heads)
    assert abs(layers[0]["bn"]["eps"] - 1e-5) < 1e-9


def test_parameters_equal_the_checkpoint():
    m = _model().load_onnx(ONNX)
    ref = _model()
    ref.load_state_dict(torch.load(PTH, map_location="cpu"))
    a, b = m.state_dict(), ref.state_dict()
    assert all(torch.equal(a[k], b[k]) for k in b if not k.endswith("num_batches_tracked"))     # the same tensors, bit for bit
    # folded: identical up to the export's float32 epsilon (9.99999975e-06 instead of 1e-5)
    fa, fb = m.folded_blob(), ref.folded_blob()
    assert float(np.abs(fa - fb).max()) <= 1e-7 * float(np.abs(fb).max())
    # a checkpoint loaded afterwards brings the reference's eps back
    m.load_state_dict(torch.load(PTH, map_location="cpu"))
    assert np.array_equal(m.folded_blob(), fb)


def test_rejects_foreign_files(tmp_path):
    bad = tmp_path / "x.onnx"
    bad.write_bytes(b"\x00\x01\x02 not a protobuf")
    with pytest.raises(yf.YfError):
        onnx_loader.read_onnx(str(bad))
    with pytest.raises(yf.YfError):                                     # a network with another head size
        yf.YoloFastest({"num_cls": 80, "input_channel": 1, "num_anchors": 3}).load_onnx(ONNX)
