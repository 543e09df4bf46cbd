"""The head kernel against the oracle: decode lists, NMS keep indices bit-exact on identical boxes, the
80-class dense-box stress case, the validation flavour, and the edge cases (empty, ties, zero-area boxes,
truncation)."""
import ctypes as C

import numpy as np
import pytest
import torch

import yolo_fastest_b200 as yf
from oracle import yolo_oracle as O
from yolo_fastest_b200 import _lib

from conftest import rows_equal

pytestmark = pytest.mark.gpu


def _pp(res, conf=None):
    io = yf.config_for(res)["io_params"]
    return io, yf.YOLO_post_process(io["conf_thre"] if conf is None else conf, io["nms_thre"], io["num_anchors"], io["num_cls"],
                                    io["anchors"], io["input_shape"])


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_decode_and_nms_on_golden_heads(gold, res):
    g = gold.res[res]
    io, pp = _pp(res)
    n = len(g["head_large"])
    pred = (torch.from_numpy(g["head_large"]).cuda(), torch.from_numpy(g["head_small"]).cuda())
    dec = pp.decode_box_batch(pred)
    kept = pp.postprocess_batch(pred)
    for i in range(n):
        rows_equal(g["decoded_%02d" % i], dec[i])
        rows_equal(g["kept_%02d" % i], kept[i])
    rows_equal(g["decoded_00"], pp.decode_box(pred))                      # reference API: batch element 0


@pytest.mark.parametrize("conf", [0.05, 0.001])
def test_low_threshold_many_survivors(gold, conf):
    """Hundreds to thousands of survivors per image with the 3-class heads: large per-class segments."""
    g = gold.res["256x320"]
    io, pp = _pp("256x320", conf)
    pred = (torch.from_numpy(g["head_large"][:3]).cuda(), torch.from_numpy(g["head_small"][:3]).cuda())
    got = pp.postprocess_batch(pred)
    for b in range(3):
        want = O.detect_postprocess((g["head_large"], g["head_small"]), io["anchors"], io["input_shape"], conf, io["nms_thre"],
                                    3, 3, batch_index=b)
        rows_equal(want, got[b])


def test_stress_80_classes(gold):
    g = gold.stress
    pp = yf.YOLO_post_process(0.001, 0.2, 3, 80, yf.COCO_ANCHORS, [416, 416, 1])
    pred = (torch.from_numpy(g["head_large"]).cuda(), torch.from_numpy(g["head_small"]).cuda())
    dec = pp.decode_box_batch(pred)
    assert [len(d) for d in dec] == [2535, 2535]                         # every candidate survives conf 0.001
    got = pp.postprocess_batch(pred)
    for b in range(2):
        rows_equal(g["kept_%02d" % b], got[b])


def test_nms_keep_indices_bit_exact():
    """Identical integer boxes in, identical keep list out (detect.py:69-84), including zero-area boxes
    (0/0 -> NaN -> kept) and heavy overlap."""
    rng = np.random.default_rng(0)
    pp = yf.YOLO_post_process(0.5, 0.2, 3, 3, yf.config_for("256x320")["io_params"]["anchors"], [256, 320, 1])
    for n, span, thr in ((1, 50, 0.2), (2, 10, 0.2), (33, 60, 0.2), (500, 300, 0.2), (1500, 200, 0.45), (64, 8, 0.0), (300, 40, 0.7)):
        xy = rng.integers(-20, span, size=(n, 2))
        wh = rng.integers(0, max(2, span // 3), size=(n, 2))
        conf = np.sort(rng.random(n))[::-1]
        rows = [[int(x), int(y), int(x + w), int(y + h), float(c), 0.5, 0] for (x, y), (w, h), c in zip(xy, wh, conf)]
        pp.nms_thres = thr
        want = O.nms([list(r) for r in rows], thr)
        got = pp.non_maxium_supression(rows)
        assert got == want, "n=%d thr=%g" % (n, thr)
    assert pp.non_maxium_supression([]) == []


def test_nms_f32_keep_indices_bit_exact():
    """Validation flavour IoU (+1, +1e-16, fp32) on identical boxes (general.py:29-52,121-136)."""
    rng = np.random.default_rng(1)
    l = yf.lib()
    ctx = _lib.Ctx(0, 1, 3, 3, 1, 64, 64)
    for n, thr in ((1, 0.2), (40, 0.2), (700, 0.4), (257, 0.05)):
        xy = rng.random((n, 2)).astype(np.float32) * 200
        wh = rng.random((n, 2)).astype(np.float32) * 60
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        conf = np.sort(rng.random(n).astype(np.float32))[::-1].copy()
        det = torch.from_numpy(np.concatenate([boxes, conf[:, None], np.ones((n, 1), np.float32), np.zeros((n, 1), np.float32)], 1))
        want = []
        alive = list(range(n))
        while alive:                                                    # general.py:127-136 on indices
            k = alive.pop(0)
            want.append(k)
            if not alive:
                break
            ious = O.bbox_iou(det[k:k + 1], det[alive])
            alive = [a for a, v in zip(alive, (ious < thr).tolist()) if v]
        d_boxes = torch.from_numpy(boxes).cuda()
        keep = torch.empty(n, dtype=torch.int32, device="cuda")
        nk = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(l.yf_nms_sorted_f32(ctx.handle, d_boxes.data_ptr(), n, C.c_float(thr), keep.data_ptr(), nk.data_ptr(), None), ctx.handle)
        torch.cuda.synchronize()
        assert keep[:int(nk.item())].cpu().tolist() == want, "n=%d" % n
    ctx.close()


def _synthetic_heads(B, nc, h, w, seed, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    A = 3
    hl = torch.randn((B, A * (5 + nc), h, w), generator=g) * scale
    hs = torch.randn((B, A * (5 + nc), h // 2, w // 2), generator=g) * scale
    return hl, hs


def test_synthetic_heads_batch(gold):
    """Random logits at batch 16: ~half of all candidates survive, ties and truncation included."""
    io, pp = _pp("256x320")
    hl, hs = _synthetic_heads(16, 3, 16, 20, 5)
    hl[:, 4] = hl[:, 4].round()                 # exact conf ties between candidates -> stable order matters
    hl[3] = -30.0                               # an image with no survivor on the large head
    hs[3] = -30.0
    got = pp.postprocess_batch((hl.cuda(), hs.cuda()))
    dec = pp.decode_box_batch((hl.cuda(), hs.cuda()))
    for b in range(16):
        want_dec = O.decode_box((hl, hs), io["anchors"], io["input_shape"], io["conf_thre"], 3, 3, batch_index=b)
        rows_equal(want_dec, dec[b], conf_tol=1e-12)
        want = O.detect_postprocess((hl, hs), io["anchors"], io["input_shape"], io["conf_thre"], io["nms_thre"], 3, 3, batch_index=b)
        rows_equal(want, got[b], conf_tol=1e-12)
    assert got[3] == [] and dec[3] == []
    # no silent truncation: an explicit capacity that some image exceeds raises (the reference's lists have no cap); the default
    # call above started from 256 slots per image and was repeated with the exact size (these images hold ~600 candidates each)
    assert max(len(d) for d in dec) > pp.DEFAULT_CAP
    with pytest.raises(yf.YfError):
        pp.postprocess_batch((hl.cuda(), hs.cuda()), max_det=5)
    # the C ABI itself reports the true count and fills the slab with the first max_det rows in output order
    ctx = pp._context(torch.device("cuda:0"), 16)
    out = torch.empty((16, 5, 56), dtype=torch.uint8, device="cuda")
    cnt = torch.empty(16, dtype=torch.int32, device="cuda")
    st = torch.empty(16, dtype=torch.int32, device="cuda")
    p = pp._params(_lib.MODE_DETECT, 5)
    a, b_ = hl.cuda().contiguous(), hs.cuda().contiguous()
    _lib.check(_lib.lib().yf_postprocess(ctx.handle, a.data_ptr(), b_.data_ptr(), 16, 16, 20, 8, 10, C.byref(p), out.data_ptr(), cnt.data_ptr(),
                                         st.data_ptr(), None), ctx.handle)
    torch.cuda.synchronize()
    slab = out.cpu().numpy().view(_lib.DET_DTYPE).reshape(16, 5)
    from yolo_fastest_b200.detector import _rows_from_dets
    for b in range(16):
        assert int(cnt[b]) == len(got[b])
        rows_equal(got[b][:5], _rows_from_dets(slab[b, :min(5, len(got[b]))]))


def test_validation_flavour_from_heads_and_rows(gold):
    """YOLOLossV3(targets=None) + non_max_suppression (validate.py:38-44) — fp32: values within 1e-5 relative,
    same rows in the same order."""
    for res in ("256x320", "512x640"):
        g = gold.res[res]
        io = yf.config_for(res)["io_params"]
        heads = (torch.from_numpy(g["head_large"]), torch.from_numpy(g["head_small"]))
        want_v = torch.cat([O.val_decode(heads[h], io["anchors"][h], io["num_cls"], io["input_shape"]) for h in range(2)], 1)
        losses = [yf.YOLOLossV3(io["anchors"][h], io["num_cls"], io["input_shape"], torch.device("cuda")) for h in range(2)]
        got_v = torch.cat([losses[h](heads[h].cuda()) for h in range(2)], 1)
        assert got_v.shape == want_v.shape
        assert torch.allclose(got_v.cpu(), want_v, rtol=1e-5, atol=1e-6)
        # NMS fed the oracle's own rows: identical selection, identical values
        got = yf.non_max_suppression(want_v.cuda(), io["num_cls"], io["conf_thre"], io["nms_thre"])
        want = O.val_nms(want_v, io["num_cls"], io["conf_thre"], io["nms_thre"])
        for b in range(len(want)):
            assert (got[b] is None) == (want[b] is None)
            if want[b] is not None:
                assert torch.equal(got[b].cpu(), want[b]), "image %d" % b
            if b < len(g["val_%02d" % 0]) * 0 + 5:
                w = g["val_%02d" % b]
                assert (len(w) == 0) == (want[b] is None)
        # low threshold: many survivors per class
        got = yf.non_max_suppression(want_v[:2].cuda(), io["num_cls"], 0.01, 0.3)
        want = O.val_nms(want_v[:2], io["num_cls"], 0.01, 0.3)
        for b in range(2):
            assert torch.equal(got[b].cpu(), want[b])


def test_fused_validate_mode_matches_rows_mode(gold):
    """yf_postprocess in YF_MODE_VALIDATE (decode fused with NMS) == yf_val_decode -> yf_val_nms."""
    g = gold.res["256x320"]
    io, pp = _pp("256x320", 0.05)
    heads = (torch.from_numpy(g["head_large"][:4]).cuda(), torch.from_numpy(g["head_small"][:4]).cuda())
    fused, _, _ = pp._run(heads, nms=True, mode=_lib.MODE_VALIDATE)
    losses = [yf.YOLOLossV3(io["anchors"][h], 3, io["input_shape"], torch.device("cuda")) for h in range(2)]
    rows = torch.cat([losses[h](heads[h]) for h in range(2)], 1)
    two = yf.non_max_suppression(rows, 3, 0.05, io["nms_thre"], raw=True)
    for b in range(4):
        assert len(fused[b]) == len(two[b]) > 0
        for f in ("x1", "y1", "x2", "y2", "conf", "cls_score", "cls", "src"):
            assert np.array_equal(fused[b][f], two[b][f]), f


def test_status_flags():
    io, pp = _pp("256x320")
    hl, hs = _synthetic_heads(2, 3, 16, 20, 9, scale=0.5)
    hl[0, 4, 0, 0] = 5.0
    hl[0, 2, 0, 0] = 30.0                       # exp(30) * anchor >> 2^25
    hl[1, 4, 1, 1] = float("nan")
    _, _, status = pp._run((hl.cuda(), hs.cuda()), nms=True, check_status=False)
    assert status[0] & 1 and status[1] & 2
    with pytest.raises(yf.YfError):
        pp.postprocess_batch((hl.cuda(), hs.cuda()))
    with pytest.raises(yf.YfError):                 # the decode-only calls check the domain flag too
        pp.decode_box_batch((hl.cuda(), hs.cuda()))
