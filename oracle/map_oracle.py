"""CPU oracle of the reference's mAP computation.  TEST INFRASTRUCTURE ONLY (same rules as yolo_oracle.py).

Restates `Validation.get_mAP` / `__calculate_AP` / `__recover_targets` (src/model_training/validate.py:27-141) on top of
`yolo_oracle.val_decode` / `val_nms`, including the reference's quirk that a match is stored as `np.array([conf, 'TP'])` — a STRING
array, so the per-class sort at :76-77 compares the decimal strings of the confidences.  Pinned against the reference's own class
by tests/golden/make_golden_map.py (bit-identical APs on the shipped images with synthetic ground truth)."""
import numpy as np
import torch

from . import yolo_oracle as O


def recover_targets(targets, input_shape):
    """validate.py:132-141: normalised (cx, cy, w, h) -> corner boxes in network-input pixels, in a copy."""
    t = targets.clone().float()
    t[:, :, (0, 2)] = t[:, :, (0, 2)] * input_shape[1]
    t[:, :, (1, 3)] = t[:, :, (1, 3)] * input_shape[0]
    xy, wh = t[:, :, 0:2].clone(), t[:, :, 2:4].clone()
    t[:, :, 0:2] = xy - wh / 2
    t[:, :, 2:4] = xy + wh / 2
    return t


def calculate_ap(match_list, target_num):
    """validate.py:91-123 for one class; match_list entries are (conf string, 'TP' | 'FP') already sorted."""
    pr = []
    for i in range(len(match_list)):
        tp = sum(1 for m in match_list[:i + 1] if m[1] == "TP")
        fp = i + 1 - tp
        fn = target_num - tp
        precision, recall = tp / (tp + fp), tp / (tp + fn)
        if i > 0 and recall == pr[-1][1]:
            if precision > pr[-1][0]:
                pr[-1][0] = precision
        else:
            pr.append([precision, recall])
    ap, prev = 0.0, 0.0
    for i in range(len(pr)):
        ap += (pr[i][1] - prev) * max(p[0] for p in pr[i:])
        prev = pr[i][1]
    return ap


def get_map(batches, num_cls, input_shape, iou_thres=0.5):
    """batches: iterable of (per-image NMS outputs as val_nms returns them, targets [B, max_boxes, 6]) -> (mAP, [AP per class],
    [targets per class])."""
    target_num = [0.0] * num_cls
    match = [[] for _ in range(num_cls)]
    for output, targets in batches:
        targets = recover_targets(targets, input_shape)
        for img_id, img_pred in enumerate(output):
            img_target = targets[img_id]
            img_target = img_target[img_target[:, 5] > 1]
            for t in img_target:
                target_num[int(t[4])] += 1
            if img_pred is None:
                continue
            for c in img_pred[:, 6].unique():
                target_c = img_target[img_target[:, 4] == c]
                for t in img_pred[img_pred[:, 6] == c]:
                    hit = None
                    if target_c.size(0):
                        ious = O.bbox_iou(t.unsqueeze(0), target_c)
                        idx = (ious > iou_thres).nonzero()
                        hit = int(idx[0]) if len(idx) else None
                    if hit is not None:
                        target_c = torch.cat((target_c[:hit], target_c[hit + 1:]))
                    match[int(c)].append(np.array([t[4], "TP" if hit is not None else "FP"]))     # a string array, as in the reference
    aps = []
    for c in range(num_cls):
        match[c].sort(key=lambda x: x[0], reverse=True)
        aps.append(calculate_ap(match[c], target_num[c]) if target_num[c] else 0.0)
    return float(sum(aps) / num_cls), [float(a) for a in aps], target_num
