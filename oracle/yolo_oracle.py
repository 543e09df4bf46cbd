"""CPU oracle for the YOLO-Fastest detection hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``yolo_fastest_b200/``
imports it, and the product path raises when its CUDA library is missing
instead of falling back to anything here.

It restates, on the CPU, the algorithm of the reference path (citations are
relative to the reference repo root):

* ``forward``            <- src/model_training/model/yolo_fastest.py:16-66,150-218
* ``forward_lite``       <- src/model_training/model/yolo_fastest.py:321-372 (YoloFastest_lite.forward)
* ``decode_box``         <- src/detect.py:23-25,41-67
* ``cal_iou`` / ``nms``  <- src/detect.py:27-39,69-84
* ``detect_postprocess`` <- src/detect.py:155-169  (class split, stable sort, per-class NMS)
* ``adjust_coord``       <- src/detect.py:131-139
* ``val_decode``         <- src/model_training/loss/yolo_loss.py:48-68,98-141 (targets=None branch)
* ``bbox_iou`` / ``val_nms`` <- src/model_training/utils/general.py:29-52,87-143

The arithmetic of ``forward`` lives in a third-party dependency of the reference
(PyTorch ATen / oneDNN CPU kernels, pinned by the reference at pytorch 1.2/1.4,
run here at 2.11); the restatement issues the same ATen ops through
``torch.nn.functional`` so it is bit-identical to the reference ``nn.Module``.

Parity pin: the reference repo holds no tests or golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is pinned instead against outputs of the
reference itself, imported in the build container from /root/reference by
``tests/golden/make_golden.py`` (which asserts oracle == reference bit for bit
and freezes the vectors under ``tests/golden/``), and against the only facts
the reference publishes: the per-image "detect finished"/"no targets" flags of
``test_result/*/笔记本cpu(python)_test_result/cpu-test.log``.

Two behaviours the reference leaves as Python exceptions are defined here the
way its C++ twin behaves (src/model_deployment/ncnn_deploy/src/YOLO_ncnn.cpp:212,221-234):
union == 0 gives IoU = NaN which never suppresses; |logit| > 709 is outside the
tested domain (``decode_box`` raises OverflowError exactly like the reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, untouched by the reference (yolo_fastest.py:12-13,220-231)

# layers built by conv_norm (no activation): yolo_fastest.py:82,89,97,104,114,134,136,144,146 and BasicResBlock.conv3 (:58)
_LINEAR = {"conv1_4", "conv2_1", "conv3_1", "conv3_4", "conv4_1", "conv5_4", "conv5_6", "conv4_1_3", "conv4_1_5"}
# stride-2 convolutions: yolo_fastest.py:78,87,95,112,122
_STRIDE2 = {"conv0", "conv1_9", "conv2_3", "conv3_6", "conv4_3"}


def _conv_bn(sd, name, x, relu):
    """conv_norm_relu / conv_norm (yolo_fastest.py:16-39): bias-free conv, pad (k-1)//2, eval-mode BN, optional ReLU."""
    w = sd[name + ".0.weight"]
    k = w.shape[-1]
    groups = x.shape[1] // w.shape[1]
    stride = 2 if name.split(".")[0] in _STRIDE2 else 1
    x = F.conv2d(x, w, None, stride=stride, padding=(k - 1) // 2, groups=groups)
    x = F.batch_norm(x, sd[name + ".1.running_mean"], sd[name + ".1.running_var"],
                     sd[name + ".1.weight"], sd[name + ".1.bias"], False, 0.1, BN_EPS)
    return F.relu(x) if relu else x


def _layer(sd, name, x):
    return _conv_bn(sd, name, x, name not in _LINEAR)


def _res(sd, name, x):
    """BasicResBlock.forward (yolo_fastest.py:60-66): 1x1+ReLU, dw3x3+ReLU, 1x1 linear, add, no ReLU after."""
    y = _conv_bn(sd, name + ".conv1", x, True)
    y = _conv_bn(sd, name + ".conv2", y, True)
    y = _conv_bn(sd, name + ".conv3", y, False)
    return y + x


_TRUNK = ["conv0", "conv1_2", "conv1_3", "conv1_4", "res1_1", "conv1_8", "conv1_9", "conv2_1", "res2_1", "res2_2",
          "conv2_2", "conv2_3", "conv3_1", "res3_1", "res3_2", "conv3_2", "conv3_3", "conv3_4", "res3_3", "res3_4",
          "res3_5", "res3_6", "conv3_5", "conv3_6", "conv4_1", "res4_1", "res4_2", "res4_3", "res4_4"]
_TRUNK5 = ["conv4_3", "conv5_1", "res5_1", "res5_2", "res5_3", "res5_4", "res5_5"]


def forward(sd, x, taps=None):
    """YoloFastest.forward (yolo_fastest.py:150-218) -> (head_large, head_small).

    ``sd`` is a state_dict as shipped in models/pytorch/*.pth, ``x`` is [B, C, H, W] fp32.
    ``taps`` (optional dict) receives named intermediate tensors for per-group parity checks.
    """
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    with torch.no_grad():
        for name in _TRUNK:
            x = _res(sd, name, x) if name.startswith("res") else _layer(sd, name, x)
            tap(name, x)
        conv4_2 = tap("conv4_2", _layer(sd, "conv4_2", x))
        x = conv4_2
        for name in _TRUNK5:
            x = _res(sd, name, x) if name.startswith("res") else _layer(sd, name, x)
            tap(name, x)
        conv5_2 = tap("conv5_2", _layer(sd, "conv5_2", x))
        x = conv5_2
        for name in ("conv5_3", "conv5_4", "conv5_5", "conv5_6"):
            x = tap(name, _layer(sd, name, x))
        head_small = F.conv2d(x, sd["head_5.weight"], sd["head_5.bias"])                      # :138,205
        # deconv_norm_relu (yolo_fastest.py:42-48,208): ConvTranspose2d k2 s2 p0, BN, ReLU
        up = F.conv_transpose2d(conv5_2, sd["deconv5_1.0.weight"], None, stride=2)
        up = F.relu(F.batch_norm(up, sd["deconv5_1.1.running_mean"], sd["deconv5_1.1.running_var"],
                                 sd["deconv5_1.1.weight"], sd["deconv5_1.1.bias"], False, 0.1, BN_EPS))
        tap("deconv5_1", up)
        x = torch.cat((conv4_2, up), 1)                                                       # :209
        for name in ("conv4_1_1", "conv4_1_2", "conv4_1_3", "conv4_1_4", "conv4_1_5"):
            x = tap(name, _layer(sd, name, x))
        head_large = F.conv2d(x, sd["head_4.weight"], sd["head_4.bias"])                      # :148,216
    return head_large, head_small


def forward_lite(sd, x, taps=None):
    """YoloFastest_lite.forward (yolo_fastest.py:321-372) -> head_5. Same layers as `forward` with conv3_3 SKIPPED (the lite forward
    goes conv3_2 -> conv3_4, :335-337) and nothing after head_5 (:365-372)."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    with torch.no_grad():
        for name in _TRUNK + ["conv4_2"] + _TRUNK5 + ["conv5_2", "conv5_3", "conv5_4", "conv5_5", "conv5_6"]:
            if name == "conv3_3":
                continue
            x = _res(sd, name, x) if name.startswith("res") else _layer(sd, name, x)
            tap(name, x)
        return F.conv2d(x, sd["head_5.weight"], sd["head_5.bias"])


def lite_state_dict(sd, num_cls=3, num_anchors=3):
    """A YoloFastest_lite state_dict derived from a shipped YoloFastest checkpoint: every layer is shape-compatible except the two
    head convs, whose (num_anchors * num_cls) * (5 + num_cls) outputs (yolo_fastest.py:240-241) are the shipped head rows repeated
    with deterministic per-copy gains, so the network stays trained-like. The reference ships no lite checkpoint."""
    out = dict(sd)
    nout = num_anchors * num_cls * (5 + num_cls)
    for h in ("head_5", "head_4"):
        w, b = sd[h + ".weight"], sd[h + ".bias"]
        reps = -(-nout // w.shape[0])
        gains = torch.linspace(1.0, 0.5, reps).repeat_interleave(w.shape[0])[:nout]
        out[h + ".weight"] = (w.repeat(reps, 1, 1, 1)[:nout] * gains.view(-1, 1, 1, 1)).contiguous()
        out[h + ".bias"] = (b.repeat(reps)[:nout] * gains).contiguous()
    return out


# ---------------------------------------------------------------------------------------------
# detect.py flavour (float64 math, integer boxes)
# ---------------------------------------------------------------------------------------------

def sigmoid(x):
    """detect.py:23-25 — float64; raises OverflowError for x < -709.78 like the reference."""
    return 1. / (1. + math.exp(-x))


def decode_box(pred, anchors, input_shape, conf_thres, num_anchors, num_cls, batch_index=0, with_src=False):
    """YOLO_post_process.decode_box (detect.py:41-67).

    ``pred`` = (head_large, head_small) as torch tensors or numpy arrays [B, A*(5+nc), h, w];
    the reference reads batch element 0 only (detect.py:46) — ``batch_index`` generalises that.
    Rows are [x1, y1, x2, y2, conf, cls_score, cls_index] with Python-int coordinates
    (round() = half-to-even), float64 conf/cls_score.  With ``with_src`` an 8th entry holds the
    running candidate index over (head, anchor, row, col), the order the reference visits them.
    """
    bbox_attrs = 5 + num_cls
    out = []
    base = 0
    for head, pred_head in enumerate(pred):
        if isinstance(pred_head, torch.Tensor):
            pred_head = pred_head.detach().cpu().numpy()
        pred_head = pred_head[batch_index]
        in_h, in_w = pred_head.shape[1], pred_head.shape[2]
        scale_h = input_shape[0] / in_h
        scale_w = input_shape[1] / in_w
        anc = anchors[head]
        p = pred_head.reshape((num_anchors, bbox_attrs, in_h, in_w))
        for pp in range(num_anchors):
            for i in range(in_h):
                for j in range(in_w):
                    conf = sigmoid(p[pp, 4, i, j])
                    if conf > conf_thres:                                   # strict (detect.py:58)
                        cls_index = np.argmax(p[pp, 5:, i, j])              # first maximum
                        cls_score = sigmoid(np.max(p[pp, 5:, i, j]))
                        x = (j + sigmoid(p[pp, 0, i, j])) * scale_w
                        y = (i + sigmoid(p[pp, 1, i, j])) * scale_h
                        w = math.exp(p[pp, 2, i, j]) * anc[pp][0]
                        h = math.exp(p[pp, 3, i, j]) * anc[pp][1]
                        row = [round(x - w / 2), round(y - h / 2), round(x + w / 2), round(y + h / 2),
                               conf, cls_score, cls_index]
                        if with_src:
                            row.append(base + (pp * in_h + i) * in_w + j)
                        out.append(row)
        base += num_anchors * in_h * in_w
    return out


def cal_iou(box_1, box_2):
    """YOLO_post_process.__cal_iou (detect.py:27-39): no +1, exact integer areas, float64 ratio.

    The reference raises ZeroDivisionError when both boxes are empty; here that is NaN, which
    the ``>`` test in ``nms`` treats as "not suppressed" (YOLO_ncnn.cpp:212,221-234).
    """
    inter_area = 0
    inter_w = min(box_1[2], box_2[2]) - max(box_1[0], box_2[0])
    inter_h = min(box_1[3], box_2[3]) - max(box_1[1], box_2[1])
    if inter_w > 0 and inter_h > 0:
        inter_area = inter_h * inter_w
    union_area = (box_1[2] - box_1[0]) * (box_1[3] - box_1[1]) + \
                 (box_2[2] - box_2[0]) * (box_2[3] - box_2[1]) - inter_area
    if union_area == 0:
        return float("nan")
    return inter_area / union_area


def nms(bbox_list, nms_thres):
    """YOLO_post_process.non_maxium_supression (detect.py:69-84) on a conf-descending list.

    Greedy: keep the head of the list, drop every later box whose IoU with it is > nms_thres
    (strict).  Written with an alive mask instead of list.pop — same result, O(n^2) not O(n^3).
    """
    n = len(bbox_list)
    alive = [True] * n
    results = []
    for a in range(n):
        if not alive[a]:
            continue
        results.append(bbox_list[a])
        for b in range(a + 1, n):
            if alive[b] and cal_iou(bbox_list[b], bbox_list[a]) > nms_thres:
                alive[b] = False
    return results


def detect_postprocess(pred, anchors, input_shape, conf_thres, nms_thres, num_anchors, num_cls,
                       batch_index=0, with_src=False):
    """decode -> split by class -> stable sort by conf desc -> per-class NMS -> concat (detect.py:155-169)."""
    boxes = decode_box(pred, anchors, input_shape, conf_thres, num_anchors, num_cls, batch_index, with_src)
    per_cls = [[] for _ in range(num_cls)]
    for b in boxes:
        per_cls[int(b[6])].append(b)
    out = []
    for c in range(num_cls):
        if not per_cls[c]:
            continue
        per_cls[c].sort(key=lambda item: item[4], reverse=True)       # list.sort is stable (detect.py:167)
        out.extend(nms(per_cls[c], nms_thres))
    return out


def adjust_coord(rows, input_shape, origin_img_shape):
    """Detect_YOLO.__adjust_coord (detect.py:131-139): round(coord * origin/input), in place."""
    scale_h = origin_img_shape[0] / input_shape[0]
    scale_w = origin_img_shape[1] / input_shape[1]
    for r in rows:
        r[0] = round(r[0] * scale_w)
        r[2] = round(r[2] * scale_w)
        r[1] = round(r[1] * scale_h)
        r[3] = round(r[3] * scale_h)


def preprocess_gray(u8_hw):
    """Numeric tail of Detect_YOLO.__pre_process (detect.py:118-127) for a 1-channel image:
    uint8 [H, W] -> fp32 [1, 1, H, W] = (x - 128) / 255.  (imread/cvtColor/resize stay with cv2.)"""
    img = torch.from_numpy(np.ascontiguousarray(u8_hw)[None]).float()
    img = (img - 128.0) / 255.0
    return img.unsqueeze(0)


# ---------------------------------------------------------------------------------------------
# validation flavour (batched fp32 torch, float boxes)
# ---------------------------------------------------------------------------------------------

def val_decode(head, anchors, num_cls, input_shape):
    """YOLOLossV3.forward(input, targets=None) (yolo_loss.py:48-68,98-141) -> [B, A*h*w, 5+nc] fp32.

    Rows ordered (anchor, row, col); columns (cx, cy, w, h) in network-input pixels, sigmoid conf,
    sigmoid per-class scores.  The reference hard-codes 3 anchors at :109-110; this takes len(anchors).
    """
    bs, _, in_h, in_w = head.shape
    na = len(anchors)
    attrs = 5 + num_cls
    stride_h = input_shape[0] / in_h
    stride_w = input_shape[1] / in_w
    scaled = [(a_w / stride_w, a_h / stride_h) for a_w, a_h in anchors]
    p = head.view(bs, na, attrs, in_h, in_w).permute(0, 1, 3, 4, 2).contiguous()
    x = torch.sigmoid(p[..., 0])
    y = torch.sigmoid(p[..., 1])
    w = p[..., 2]
    h = p[..., 3]
    conf = torch.sigmoid(p[..., 4])
    pred_cls = torch.sigmoid(p[..., 5:])
    grid_x = torch.arange(in_w).repeat(bs, na, in_h, 1).float()
    grid_y = torch.arange(in_h).repeat(bs, na, in_w, 1).permute(0, 1, 3, 2).float()
    anchor_w = torch.tensor([s[0] for s in scaled], dtype=torch.float32).view(1, na, 1, 1).expand(bs, na, in_h, in_w)
    anchor_h = torch.tensor([s[1] for s in scaled], dtype=torch.float32).view(1, na, 1, 1).expand(bs, na, in_h, in_w)
    boxes = torch.empty(p[..., :4].shape, dtype=torch.float32)
    boxes[..., 0] = x + grid_x
    boxes[..., 1] = y + grid_y
    boxes[..., 2] = torch.exp(w) * anchor_w
    boxes[..., 3] = torch.exp(h) * anchor_h
    scale = torch.tensor([stride_w, stride_h, stride_w, stride_h], dtype=torch.float32)
    return torch.cat((boxes.view(bs, -1, 4) * scale, conf.view(bs, -1, 1), pred_cls.view(bs, -1, num_cls)), -1)


def bbox_iou(box1, box2):
    """bbox_iou (general.py:29-52), x1y1x2y2 branch: +1 pixel widths, +1e-16 in the denominator, fp32."""
    b1_x1, b1_y1, b1_x2, b1_y2 = box1[:, 0], box1[:, 1], box1[:, 2], box1[:, 3]
    b2_x1, b2_y1, b2_x2, b2_y2 = box2[:, 0], box2[:, 1], box2[:, 2], box2[:, 3]
    ix1 = torch.max(b1_x1, b2_x1)
    iy1 = torch.max(b1_y1, b2_y1)
    ix2 = torch.min(b1_x2, b2_x2)
    iy2 = torch.min(b1_y2, b2_y2)
    inter = torch.clamp(ix2 - ix1 + 1, min=0) * torch.clamp(iy2 - iy1 + 1, min=0)
    a1 = (b1_x2 - b1_x1 + 1) * (b1_y2 - b1_y1 + 1)
    a2 = (b2_x2 - b2_x1 + 1) * (b2_y2 - b2_y1 + 1)
    return inter / (a1 + a2 - inter + 1e-16)


def val_nms(prediction, num_classes, conf_thres=0.5, nms_thres=0.4):
    """non_max_suppression (general.py:87-143) -> list (len B) of [n, 7] fp32 tensors or None.

    conf >= conf_thres; class = argmax of the class scores; classes ascending; per class sorted
    by conf descending; keep a box iff its IoU(+1) with every earlier kept box is < nms_thres.
    The reference's torch.sort is unstable; ties are broken here by candidate order (stable),
    which is one of the orders the reference may produce.  Does not mutate ``prediction``
    (the reference overwrites its xywh with xyxy at :95).
    """
    pred = prediction.clone()
    corner = pred.new(pred.shape)
    corner[:, :, 0] = pred[:, :, 0] - pred[:, :, 2] / 2
    corner[:, :, 1] = pred[:, :, 1] - pred[:, :, 3] / 2
    corner[:, :, 2] = pred[:, :, 0] + pred[:, :, 2] / 2
    corner[:, :, 3] = pred[:, :, 1] + pred[:, :, 3] / 2
    pred[:, :, :4] = corner[:, :, :4]
    output = [None for _ in range(len(pred))]
    for image_i, image_pred in enumerate(pred):
        image_pred = image_pred[image_pred[:, 4] >= conf_thres]
        if not image_pred.size(0):
            continue
        class_conf, class_pred = torch.max(image_pred[:, 5:5 + num_classes], dim=1, keepdim=True)
        det = torch.cat((image_pred[:, :5], class_conf.float(), class_pred.float()), 1)
        for c in det[:, -1].unique():
            dc = det[det[:, 6] == c]
            order = torch.sort(dc[:, 4], descending=True, stable=True)[1]
            dc = dc[order]
            kept = []
            while dc.size(0):
                kept.append(dc[0].unsqueeze(0))
                if len(dc) == 1:
                    break
                ious = bbox_iou(kept[-1], dc[1:])
                dc = dc[1:][ious < nms_thres]
            kept = torch.cat(kept)
            output[image_i] = kept if output[image_i] is None else torch.cat((output[image_i], kept))
    return output
