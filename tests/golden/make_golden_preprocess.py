"""Freeze cv2's own outputs for the pre-processing parity tests (run in the build container, where cv2 is importable):
    python tests/golden/make_golden_preprocess.py
Writes tests/golden/preprocess.npz: for each case the BGR frame (or the name of a committed test image), the target size and
cv2.resize(cv2.cvtColor(frame, BGR2GRAY), (W, H)) - exactly the calls of the reference's Detect_YOLO.__pre_process
(src/detect.py:107-122)."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
out = {"cv2_version": np.array(cv2.__version__)}
cases = []
# shipped test images (committed under tests/golden/images): 512x640 frames -> the two network input sizes
for name in sorted(os.listdir(os.path.join(HERE, "images")))[:3]:
    img = cv2.imread(os.path.join(HERE, "images", name))
    for (H, W) in ((256, 320), (512, 640)):
        g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        if g.shape != (H, W):
            g = cv2.resize(g, (W, H))
        key = "img_%s_%dx%d" % (name, H, W)
        out[key] = g
        cases.append(key)
# random frames, odd sizes, both directions
for i, (Ho, Wo, H, W) in enumerate([(64, 80, 32, 32), (97, 131, 64, 96), (48, 64, 96, 128), (33, 47, 64, 64), (100, 100, 32, 96),
                                    (240, 320, 256, 320)]):
    frame = rng.integers(0, 256, (Ho, Wo, 3), dtype=np.uint8)
    out["rnd%d_bgr" % i] = frame
    out["rnd%d_out" % i] = cv2.resize(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY), (W, H))
out["image_cases"] = np.array(cases)
np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
print("wrote", os.path.join(HERE, "preprocess.npz"), os.path.getsize(os.path.join(HERE, "preprocess.npz")), "bytes")
