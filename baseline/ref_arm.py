"""The reference arm of bench.py: the UNMODIFIED reference (staged under baseline/_ref by __graft_entry__.build() in the build
container, git-ignored, shipped to the GPU box by gpurun) run on the host cores through its own public API and stock code path —
`YoloFastest(io_params)` + `load_state_dict` + `model(img)` + `YOLO_post_process.decode_box` / sort / `non_maxium_supression`
exactly as `Detect_YOLO.batch_detect` drives them (src/detect.py:87-105,141-169). Nothing of this repo's model, kernels or oracle
is on that path. `decode_box` reads batch element 0 only (src/detect.py:46), so a batch is forwarded once and decoded image by image
(BASELINE.md §3)."""
import copy
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "src", "detect.py"))


_mods = None


def load():
    """Import the staged reference modules (once). The reference uses script-relative imports and needs tensorboardX only for its
    training loop (train.py:11) — stubbed, as SURVEY.md §3.1 describes."""
    global _mods
    if _mods is None:
        sys.path.insert(0, os.path.join(REF, "src", "model_training"))
        sys.path.insert(0, os.path.join(REF, "src"))
        sys.modules.setdefault("tensorboardX", types.SimpleNamespace(SummaryWriter=object))
        import detect as ref_detect                                        # src/detect.py
        from model_training._config import config_params as ref_config    # src/model_training/_config.py
        ref_detect.device = torch.device("cpu")                            # module global read by decode_box (detect.py:44)
        _mods = (ref_detect, ref_config)
    return _mods


class RefPipeline:
    def __init__(self, res, sd):
        ref_detect, ref_config = load()
        cfg = copy.deepcopy(ref_config)
        io = cfg["io_params"]
        if res == "256x320":
            io["input_shape"], io["anchors"] = [256, 320, 1], io["anchors"][0:2]      # _config.py:5-9
        else:
            io["input_shape"], io["anchors"] = [512, 640, 1], io["anchors"][1:3]
        self.io = io
        self.model = ref_detect.YoloFastest(io).eval()                                 # detect.py:89
        self.model.load_state_dict(sd)                                                 # detect.py:90-91
        self.pp = ref_detect.YOLO_post_process(io["conf_thre"], io["nms_thre"], io["num_anchors"], io["num_cls"], io["anchors"],
                                               io["input_shape"])                      # detect.py:100-105

    def run(self, u8):
        """u8 [B, H, W] uint8 -> (forward seconds, post-process seconds, detections). Pre-processing tail as detect.py:123-124."""
        x = (u8.float().unsqueeze(1) - 128.0) / 255.0
        nc = self.io["num_cls"]
        with torch.no_grad():
            t0 = time.perf_counter()
            pred = self.model(x)                                                       # detect.py:152
            t1 = time.perf_counter()
            n_det = 0
            for b in range(x.shape[0]):
                boxes = self.pp.decode_box((pred[0][b:b + 1], pred[1][b:b + 1]))       # detect.py:155
                per_cls = [[] for _ in range(nc)]
                for bb in boxes:                                                       # detect.py:158-161
                    per_cls[bb[6]].append(bb)
                for lst in per_cls:                                                    # detect.py:162-169
                    if lst:
                        lst.sort(key=lambda v: v[4], reverse=True)
                        n_det += len(self.pp.non_maxium_supression(lst))
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1, n_det
