"""Pre-processing oracle (oracle/preprocess_oracle.py) against cv2: frozen outputs and, when cv2 is importable, cv2 itself."""
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "preprocess.npz"))


def test_oracle_matches_frozen_cv2_on_random_frames():
    i = 0
    while "rnd%d_bgr" % i in GOLD:
        frame, want = GOLD["rnd%d_bgr" % i], GOLD["rnd%d_out" % i]
        got = po.pre_process(frame, *want.shape)
        assert got.dtype == np.uint8 and np.array_equal(got, want), "case %d" % i      # bit-exact
        i += 1
    assert i >= 6


def test_oracle_matches_frozen_cv2_on_shipped_images():
    cv2 = pytest.importorskip("cv2")           # only the JPEG decode needs it; the comparison target is frozen
    for key in GOLD["image_cases"]:
        key = str(key)
        name, size = key[4:].rsplit("_", 1)
        H, W = (int(v) for v in size.split("x"))
        img = cv2.imread(os.path.join(HERE, "golden", "images", name))
        assert np.array_equal(po.pre_process(img, H, W), GOLD[key]), key


def test_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for (Ho, Wo, H, W) in [(512, 640, 256, 320), (512, 640, 512, 640), (480, 640, 416, 416), (77, 201, 64, 160), (64, 64, 128, 160),
                           (300, 400, 32, 32)]:
        frame = rng.integers(0, 256, (Ho, Wo, 3), dtype=np.uint8)
        g = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        want = g if g.shape == (H, W) else cv2.resize(g, (W, H))
        assert np.array_equal(po.pre_process(frame, H, W), want), (Ho, Wo, H, W)
    # flat and extreme frames
    for v in (0, 255):
        frame = np.full((50, 70, 3), v, np.uint8)
        assert np.array_equal(po.pre_process(frame, 32, 64), cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (64, 32)))


def test_u8_normalisation_without_table_is_the_reference_division():
    """yf_wirb.cuh: norm_u8 computes (b - 128) / 255 as q + (n - 255 q) * r with r = fl(1 / 255): for every byte the result is the
    correctly rounded fp32 quotient, i.e. exactly what the reference's `(img - 128.0) / 255.0` gives (detect.py:124).  Checked in exact
    rational arithmetic: no floating-point emulation of the fused multiply-adds is trusted."""
    from fractions import Fraction as F
    r = np.float32(1.0) / np.float32(255.0)
    for b in range(256):
        n = np.float32(b) - np.float32(128.0)
        ref = np.float32(n / np.float32(255.0))
        q = np.float32(n * r)
        res_exact = F(float(n)) - 255 * F(float(q))                    # fmaf(-255, q, n) before rounding
        res = np.float32(float(res_exact))
        assert F(float(res)) == res_exact                              # the residual is exactly representable
        x = F(float(q)) + F(float(res)) * F(float(r))                  # the value the last fmaf rounds
        lo, hi = np.nextafter(ref, np.float32(-np.inf)), np.nextafter(ref, np.float32(np.inf))
        d = abs(x - F(float(ref)))
        assert d < abs(x - F(float(lo))) and d < abs(x - F(float(hi))), b
