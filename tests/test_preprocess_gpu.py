"""GPU pre-processing (yf_preprocess_bgr / yf_detect_host_bgr) against the CPU oracle: bit-exact bytes, identical detections."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _detector(res):
    import yolo_fastest_b200 as yf
    return yf.Detect_YOLO(torch.device("cuda:0"), os.path.join(HERE, "golden", "weights", "yolo_fastest_%s.pth" % res), yf.config_for(res), None)


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_kernel_bytes_equal_oracle(res):
    from oracle import preprocess_oracle as po
    det = _detector(res)
    H, W = det.input_shape[0:2]
    rng = np.random.default_rng(3)
    # frame sizes: the dataset's (downscale by 2 or identity), odd sizes both ways, an upscale, a strong downscale
    for (B, Ho, Wo) in [(5, 512, 640), (3, 480, 640), (2, 333, 517), (2, 128, 160), (1, 1080, 1920), (4, H, W)]:
        frames = rng.integers(0, 256, (B, Ho, Wo, 3), dtype=np.uint8)
        got = det.pre_process_batch(frames).cpu().numpy()
        for b in range(B):
            want = po.pre_process(frames[b], H, W)
            assert np.array_equal(got[b], want), (res, B, Ho, Wo, b)          # bit-exact
    # extremes: saturated frames stay saturated
    for v in (0, 255):
        frames = np.full((1, 100, 200, 3), v, np.uint8)
        assert (det.pre_process_batch(frames).cpu().numpy() == v).all()


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_detect_from_frames_matches_host_preprocessing(res):
    """The 20 shipped test images through yf_detect_host_bgr == the same images pre-processed on the host by the oracle and
    sent through yf_detect_host_u8, box for box; boxes are reported in frame coordinates like batch_detect."""
    cv2 = pytest.importorskip("cv2")
    from oracle import preprocess_oracle as po
    det = _detector(res)
    H, W = det.input_shape[0:2]
    names = sorted(os.listdir(os.path.join(HERE, "golden", "images")))
    frames = np.stack([cv2.imread(os.path.join(HERE, "golden", "images", n)) for n in names])
    rows_bgr = det.detect_bgr_batch(frames)
    gray = np.stack([po.pre_process(f, H, W) for f in frames])
    rows_u8 = det.detect_batch(gray)
    for rr in rows_u8:
        if [H, W] != list(frames.shape[1:3]):
            det.adjust_coord(rr)
    assert rows_bgr == rows_u8
    assert sum(len(r) for r in rows_bgr) > 0


def test_bad_arguments_fail_loudly():
    import yolo_fastest_b200 as yf
    det = _detector("256x320")
    with pytest.raises(yf.YfError):
        det.pre_process_batch(np.zeros((2, 64, 64), np.uint8))
    with pytest.raises(yf.YfError):
        det.detect_bgr_batch(torch.zeros((1, 64, 64, 3), dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("res", ["256x320", "512x640"])
def test_batched_file_driver_matches_reference_logs(res, tmp_path):
    """batch_detect_batched writes the same files and detect / no-target pattern as batch_detect (the published logs:
    every 256x320 image has detections, at 512x640 only noCloud_2m_4359.jpg has none)."""
    import logging
    import yolo_fastest_b200 as yf
    pytest.importorskip("cv2")
    records = []

    class Hd(logging.Handler):
        def emit(self, r):
            records.append(r.getMessage())
    logger = logging.getLogger("yf-test-batched-" + res)
    logger.setLevel(logging.INFO)
    logger.addHandler(Hd())
    det = yf.Detect_YOLO(torch.device("cuda:0"), os.path.join(HERE, "golden", "weights", "yolo_fastest_%s.pth" % res), yf.config_for(res), logger)
    out = tmp_path / "out"
    out.mkdir()
    det.batch_detect_batched(os.path.join(HERE, "golden", "images"), str(out), batch_size=8)
    assert len(os.listdir(out)) == 20 and len(records) == 21
    none = [r for r in records if "no targets" in r]
    if res == "512x640":
        assert len(none) == 1 and "noCloud_2m_4359.jpg" in none[0]
    else:
        assert not none
    assert records[-1].startswith("detect avg_time:")
