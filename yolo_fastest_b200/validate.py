"""Validation driver on the B200 path: the mAP loop of the reference's `Validation.get_mAP`
(src/model_training/validate.py:27-89) and its VOC-XML label reader (src/model_training/dataloader/detect_dataset.py:63-84).

Per batch the reference runs `model(imgs)`, one `YOLOLossV3(head)` decode per scale, `torch.cat`, `non_max_suppression`
(validate.py:38-44).  Here that is ONE fused device call per batch: `yf_forward` followed by `yf_postprocess` in YF_MODE_VALIDATE
(fp32 decode of both scales + confidence filter + per-class NMS with the +1 IoU convention, csrc/yf_post.cuh) — the decoded
`[B, A*h*w, 5+nc]` tensor never exists.  What remains on the host is what the reference also does in Python on a handful of boxes per
image: matching predictions with ground truth (first target of the class with IoU > IOU_val_thre, each target at most once,
:57-74) and integrating the precision/recall list (:91-123).  `tests/test_validate_*.py` pin this against an oracle that was itself
checked against the reference's own class.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np
import torch

from . import _lib
from .detector import YOLO_post_process


def read_voc_xml(xml_path, class_names):
    """One label file -> [[cls_index, xmin, ymin, xmax, ymax], ...] (detect_dataset.py:67-80: every <object>, its <name> looked up
    in the class list, its <bndbox> corners as floats)."""
    labels = []
    for obj in ET.parse(xml_path).findall("object"):
        box = obj.find("bndbox")
        labels.append([class_names.index(obj.find("name").text)] + [float(box.find(k).text) for k in ("xmin", "ymin", "xmax", "ymax")])
    return labels


def list_voc_folder(dataset_dir, class_names):
    """`<dir>/xml/*.xml` + `<dir>/img/<same name>.jpg` (detect_dataset.py:60-84) -> [(image path, labels)] in directory order."""
    xml_dir, img_dir = os.path.join(dataset_dir, "xml"), os.path.join(dataset_dir, "img")
    return [(os.path.join(img_dir, os.path.splitext(f)[0] + ".jpg"), read_voc_xml(os.path.join(xml_dir, f), class_names))
            for f in os.listdir(xml_dir)]


def targets_tensor(labels_per_image, origin_shape, max_boxes=64):
    """Labels in origin-image pixels -> the reference's padded target tensor [B, max_boxes, 6] = (cx, cy, w, h normalised to the
    image, class, 255 flag for valid rows) — the layout `Validation.__recover_targets` expects (validate.py:132-141,51)."""
    H, W = origin_shape[0], origin_shape[1]
    t = np.zeros((len(labels_per_image), max_boxes, 6), dtype=np.float32)
    for b, labels in enumerate(labels_per_image):
        for i, (c, x1, y1, x2, y2) in enumerate(labels[:max_boxes]):
            t[b, i] = [(x1 + x2) / 2 / W, (y1 + y2) / 2 / H, (x2 - x1) / W, (y2 - y1) / H, c, 255.0]
    return torch.from_numpy(t)


def _iou_plus1(box, targets):
    """bbox_iou of one box with [n, 4] targets, x1y1x2y2, +1 pixel widths and +1e-16 (utils/general.py:29-52), in fp32"""
    f = np.float32
    ix1, iy1 = np.maximum(box[0], targets[:, 0]), np.maximum(box[1], targets[:, 1])
    ix2, iy2 = np.minimum(box[2], targets[:, 2]), np.minimum(box[3], targets[:, 3])
    inter = np.clip(ix2 - ix1 + f(1), 0, None) * np.clip(iy2 - iy1 + f(1), 0, None)
    a1 = (box[2] - box[0] + f(1)) * (box[3] - box[1] + f(1))
    a2 = (targets[:, 2] - targets[:, 0] + f(1)) * (targets[:, 3] - targets[:, 1] + f(1))
    return inter / (a1 + a2 - inter + f(1e-16))


def average_precision(matches, target_num):
    """matches: [(conf, is_tp)] of one class sorted by conf descending -> AP as validate.py:91-123 integrates it: one
    (precision, recall) point per prefix, points of equal recall merged keeping the larger precision, area = sum over points of
    (recall step) x (largest precision from this point on)."""
    pr = []
    tp = fp = 0
    for _, is_tp in matches:
        tp, fp = tp + bool(is_tp), fp + (not is_tp)
        precision, recall = tp / (tp + fp), tp / float(target_num)          # recall = TP / (TP + FN), FN = targets - TP
        if pr and recall == pr[-1][1]:
            pr[-1][0] = max(pr[-1][0], precision)
        else:
            pr.append([precision, recall])
    ap, prev = 0.0, 0.0
    for i, (_, recall) in enumerate(pr):
        ap += (recall - prev) * max(p for p, _ in pr[i:])
        prev = recall
    return ap


class Validation:
    """Drop-in for the reference class (validate.py:8-25): `Validation(params, logger, dataset, device, model_loss)`; `dataset` is
    any iterable of (imgs [B, C, H, W] float, normalised as the network expects, targets [B, max_boxes, 6]) batches — e.g. a
    torch DataLoader over the reference's DetectDataset; `model_loss` is accepted for signature compatibility and unused (the decode
    is fused into the post-processing kernel)."""

    def __init__(self, params, logger, dataset, device, model_loss=None):
        io = params["io_params"]
        self.logger, self.dataset, self.device = logger, dataset, torch.device(device)
        self.input_shape, self.num_cls, self.cls_name = io["input_shape"], io["num_cls"], io["class_names"]
        self.IOU_threshold = params.get("train_params", {}).get("IOU_val_thre", 0.5)
        self.conf_thres, self.nms_thres = io["conf_thre"], io["nms_thre"]
        self.post = YOLO_post_process(self.conf_thres, self.nms_thres, io["num_anchors"], self.num_cls, io["anchors"], self.input_shape)
        self.clear()

    def clear(self):
        self.target_num = [0] * self.num_cls
        self.match_list = [[] for _ in range(self.num_cls)]

    def predict(self, model, imgs):
        """imgs -> per image an [n, 7] float32 array (x1, y1, x2, y2, obj_conf, class_conf, class) in the reference's order (classes
        ascending, conf descending): forward + fused validate-mode post-processing on the device."""
        dets, _, _ = self.post._run(model(imgs.to(self.device).float()), nms=True, mode=_lib.MODE_VALIDATE)
        return [np.stack([d["x1"], d["y1"], d["x2"], d["y2"], d["conf"], d["cls_score"], d["cls"].astype(np.float64)], 1).astype(np.float32)
                if len(d) else np.zeros((0, 7), np.float32) for d in dets]

    def accumulate(self, preds, targets):
        """validate.py:46-74 for one batch: targets back to input pixels as corner boxes, then per image and predicted class the
        greedy match of every prediction (in list order) against the not yet matched targets of that class."""
        t = targets.numpy().astype(np.float32).copy() if isinstance(targets, torch.Tensor) else np.asarray(targets, np.float32).copy()
        in_h, in_w = np.float32(self.input_shape[0]), np.float32(self.input_shape[1])
        t[:, :, [0, 2]] *= in_w
        t[:, :, [1, 3]] *= in_h
        xy, wh = t[:, :, 0:2].copy(), t[:, :, 2:4].copy()
        t[:, :, 0:2], t[:, :, 2:4] = xy - wh / 2, xy + wh / 2              # xywh2xyxy (utils/general.py:16-25)
        for pred, tgt in zip(preds, t):
            tgt = tgt[tgt[:, 5] > 1]
            for row in tgt:
                self.target_num[int(row[4])] += 1
            for c in np.unique(pred[:, 6]) if len(pred) else []:
                left = tgt[tgt[:, 4] == c]
                for p in pred[pred[:, 6] == c]:
                    hit = np.nonzero(_iou_plus1(p, left) > self.IOU_threshold)[0] if len(left) else []
                    if len(hit):
                        left = np.delete(left, hit[0], axis=0)              # each target is matched at most once (:70)
                    self.match_list[int(c)].append((float(p[4]), len(hit) > 0))

    def get_mAP(self, model, epoch=0):
        self.clear()
        model.eval()
        with torch.no_grad():
            for imgs, targets in self.dataset:
                self.accumulate(self.predict(model, imgs), targets)
        aps = []
        if self.logger:
            self.logger.info("—————— epoch: %d validation results —————" % epoch)
        for c in range(self.num_cls):
            self.match_list[c].sort(key=lambda m: m[0], reverse=True)         # stable, like list.sort in the reference (:76-77)
            ap = average_precision(self.match_list[c], self.target_num[c]) if self.target_num[c] else 0.0
            aps.append(ap)
            if self.logger:
                self.logger.info("class: %s, target_num = %d, AP = %.3f" % (self.cls_name[c], self.target_num[c], ap))
        m = float(sum(aps) / self.num_cls)
        if self.logger:
            self.logger.info("mean AP: %.3f" % m)
            self.logger.info("——————————————————————————")
        self.APs = aps
        return m
