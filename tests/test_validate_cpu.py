"""Validation driver, host side: the mAP oracle against the numbers frozen from the reference's own Validation class
(tests/golden/make_golden_map.py), and the product's host logic (target recovery, greedy matching, AP integration, VOC-XML reader)
against that oracle — fed with the oracle's detections, so no GPU is involved."""
import os

import numpy as np
import pytest
import torch

import yolo_fastest_b200 as yf
from oracle import map_oracle as M
from oracle import yolo_oracle as O
from yolo_fastest_b200 import validate as V

BS = 5


def _oracle_batches(gold, res):
    g, io = gold.res[res], yf.config_for(res)["io_params"]
    heads = (torch.from_numpy(g["head_large"]), torch.from_numpy(g["head_small"]))
    targets = torch.from_numpy(np.load(os.path.join(gold.dir, "golden_map.npz"))["targets_" + res])
    out = []
    for i in range(0, heads[0].shape[0], BS):
        rows = torch.cat([O.val_decode(heads[h][i:i + BS], io["anchors"][h], io["num_cls"], io["input_shape"]) for h in range(2)], 1)
        out.append((O.val_nms(rows, io["num_cls"], io["conf_thre"], io["nms_thre"]), targets[i:i + BS]))
    return out, io


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_map_oracle_reproduces_the_reference(gold, res):
    gm = np.load(os.path.join(gold.dir, "golden_map.npz"))
    batches, io = _oracle_batches(gold, res)
    m, aps, _ = M.get_map(batches, io["num_cls"], io["input_shape"], 0.5)
    assert m == float(gm["map_" + res]) and aps == [float(a) for a in gm["aps_" + res]]


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_driver_host_logic_equals_the_oracle(gold, res):
    gm = np.load(os.path.join(gold.dir, "golden_map.npz"))
    batches, io = _oracle_batches(gold, res)
    cfg = yf.config_for(res)
    v = V.Validation(cfg, None, None, "cuda:0")                      # constructing it needs no GPU
    for dets, targets in batches:
        preds = [d.numpy() if d is not None else np.zeros((0, 7), np.float32) for d in dets]
        v.accumulate(preds, targets)
    aps = []
    for c in range(io["num_cls"]):
        v.match_list[c].sort(key=lambda m: m[0], reverse=True)
        aps.append(V.average_precision(v.match_list[c], v.target_num[c]) if v.target_num[c] else 0.0)
    assert np.allclose(aps, gm["aps_" + res], rtol=0, atol=1e-12) and abs(sum(aps) / 3 - float(gm["map_" + res])) < 1e-12


def test_average_precision_cases():
    assert V.average_precision([], 4) == 0.0
    assert V.average_precision([(0.9, True), (0.8, True)], 2) == 1.0
    assert V.average_precision([(0.9, False), (0.8, True)], 1) == 0.5              # precision 0, then 1/2 at recall 1
    assert abs(V.average_precision([(0.9, True), (0.8, False), (0.7, True)], 4) - (0.25 * 1.0 + 0.25 * 2 / 3)) < 1e-15


def test_voc_reader_and_targets(tmp_path):
    (tmp_path / "xml").mkdir()
    (tmp_path / "img").mkdir()
    (tmp_path / "xml" / "a.xml").write_text(
        "<annotation><object><name>%s</name><bndbox><xmin>10</xmin><ymin>20</ymin><xmax>110</xmax><ymax>70</ymax></bndbox></object>"
        "<object><name>%s</name><bndbox><xmin>0</xmin><ymin>0</ymin><xmax>64</xmax><ymax>32</ymax></bndbox></object></annotation>"
        % (yf.config_params["io_params"]["class_names"][2], yf.config_params["io_params"]["class_names"][0]))
    names = yf.config_params["io_params"]["class_names"]
    items = V.list_voc_folder(str(tmp_path), names)
    assert items == [(str(tmp_path / "img" / "a.jpg"), [[2, 10.0, 20.0, 110.0, 70.0], [0, 0.0, 0.0, 64.0, 32.0]])]
    t = V.targets_tensor([items[0][1]], [512, 640, 3])
    assert t.shape == (1, 64, 6) and t[0, 0].tolist() == [60 / 640, 45 / 512, 100 / 640, 50 / 512, 2.0, 255.0] and float(t[0, 2, 5]) == 0.0
    back = M.recover_targets(t, [512, 640, 1])[0, 0, :4].tolist()
    assert np.allclose(back, [10.0, 20.0, 110.0, 70.0], atol=1e-4)
