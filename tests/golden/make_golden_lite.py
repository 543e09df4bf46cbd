"""Pin oracle.forward_lite against the UNMODIFIED reference's YoloFastest_lite (build container only; needs /root/reference).

    python tests/golden/make_golden_lite.py

The reference ships no lite checkpoint: the state_dict is derived from the shipped 256x320 checkpoint by
oracle.yolo_oracle.lite_state_dict (head convs widened to the lite head size). Asserts oracle == reference module bit for bit on
seeded inputs at two sizes and writes tests/golden/golden_lite.npz (inputs are regenerated from the seed by the tests)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src", "model_training"))
sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))

from model.yolo_fastest import YoloFastest_lite as RefLite  # noqa: E402
from oracle import yolo_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
sd = O.lite_state_dict(torch.load(os.path.join(HERE, "weights", "yolo_fastest_256x320.pth"), map_location="cpu"))
ref = RefLite({"num_cls": 3, "input_channel": 1, "num_anchors": 3}).eval()
print(ref.load_state_dict(sd))
out = {"seed": 41}
for tag, (B, H, W) in {"a": (2, 256, 320), "b": (1, 96, 352)}.items():
    x = (torch.randint(0, 256, (B, 1, H, W), generator=torch.Generator().manual_seed(41)).float() - 128.0) / 255.0
    want = ref(x)
    got = O.forward_lite(sd, x)
    assert want.shape == (B, 72, H // 32, W // 32) and torch.equal(want, got), tag
    out["head_" + tag] = want.numpy()
    out["shape_" + tag] = np.array([B, H, W])
np.savez_compressed(os.path.join(HERE, "golden_lite.npz"), **out)
print("oracle.forward_lite == reference YoloFastest_lite bit for bit; wrote golden_lite.npz")
