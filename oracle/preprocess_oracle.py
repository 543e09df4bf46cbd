"""CPU oracle for the image half of the pre-processing.  TEST INFRASTRUCTURE ONLY (same rules as yolo_oracle.py: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it).

Restates ``Detect_YOLO.__pre_process`` of the reference (src/detect.py:107-122): ``cv2.cvtColor(BGR2GRAY)`` followed by
``cv2.resize(img, (W, H))`` (INTER_LINEAR) on uint8 frames.  The arithmetic lives in a third-party dependency that is not
under /root/reference: OpenCV (opencv-python 4.13.0 in this image; the reference does not pin a version).  For 8-bit
images OpenCV computes both steps in integers, and that published algorithm is what is restated here:

* gray (imgproc color_rgb, RGB2Gray<uchar>): Y = (B*3735 + G*19235 + R*9798 + 2^14) >> 15
* resize (imgproc resize.cpp, INTER_LINEAR, uchar): fx = (float)((dx + 0.5) * scale - 0.5), sx = floor(fx), fx -= sx;
  columns clamp sx to [0, Wo - 1] and zero the fraction at the borders, rows keep the fraction and clip the two source
  rows; coefficients are cvRound(c * 2048) (half to even); horizontal pass h = S[sx]*a0 + S[sx+1]*a1, vertical pass
  dst = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2.

Parity pin: tests/test_preprocess_cpu.py checks it bit for bit against cv2 itself (live, when cv2 is importable) and
against tests/golden/preprocess.npz, frozen from cv2 by tests/golden/make_golden_preprocess.py.
"""
import numpy as np


def bgr2gray(bgr):
    """uint8 [..., 3] (B, G, R) -> uint8 [...]"""
    v = bgr.astype(np.int64)
    return ((v[..., 0] * 3735 + v[..., 1] * 19235 + v[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _taps(dst_n, src_n, is_row):
    scale = 1.0 / (dst_n / src_n)
    d = np.arange(dst_n)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s).astype(np.float32)
    if is_row:
        s0, s1 = np.clip(s, 0, src_n - 1), np.clip(s + 1, 0, src_n - 1)
    else:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src_n - 1
        f[hi] = 0
        s[hi] = src_n - 1
        s0, s1 = s, np.minimum(s + 1, src_n - 1)
    c0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    c1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s0, s1, c0, c1


def resize_linear(gray, H, W):
    """uint8 [Ho, Wo] -> uint8 [H, W], cv2.resize(gray, (W, H)) with the default INTER_LINEAR"""
    Ho, Wo = gray.shape
    sx0, sx1, a0, a1 = _taps(W, Wo, False)
    sy0, sy1, b0, b1 = _taps(H, Ho, True)
    S = gray.astype(np.int64)
    h = S[:, sx0] * a0 + S[:, sx1] * a1
    h0, h1 = h[sy0], h[sy1]
    return ((((b0[:, None] * (h0 >> 4)) >> 16) + ((b1[:, None] * (h1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def pre_process(bgr, H, W):
    """detect.py:107-122 for one frame: uint8 [Ho, Wo, 3] -> uint8 [H, W] (the resize is skipped when the sizes agree)"""
    g = bgr2gray(bgr)
    return g if g.shape == (H, W) else resize_linear(g, H, W)
