"""Weights from the reference's ncnn deployment files: `models/ncnn/<res>/*-opt.param` + `.bin`, the BN-folded
network the reference's C++ twin runs (`src/model_deployment/ncnn_deploy/YOLO_ncnn.cpp:23-47`).

The `.param` text lists the layers in execution order; every `Convolution`, `ConvolutionDepthWise` and `Deconvolution` layer owns,
in that order in the `.bin`, a 4-byte storage tag (0 = raw fp32), `6=` weight floats and - with `5=1` - `0=` bias floats
(ncnn `ModelBinFromDataReader::load`, type 0). ncnnoptimize has already folded the BatchNorms, so the arrays are exactly the
`(w', b')` pairs `YoloFastest.folded_blob` computes from a `.pth`, in the same (forward) order; the transposed convolution is
stored `[out][in][kh][kw]` and is turned back into PyTorch's `[in][out][kh][kw]`."""
import numpy as np

from . import _lib

_CONV_TYPES = ("Convolution", "ConvolutionDepthWise", "Deconvolution", "DeconvolutionDepthWise")


def read_ncnn(param_path, bin_path):
    """-> list of dicts {type, name, num_output, kernel, stride, group, weight (flat fp32), bias (fp32 or None)} in file order"""
    with open(param_path, "r") as f:
        lines = f.read().split("\n")
    if not lines or lines[0].strip() != "7767517":
        raise _lib.YfError("%s is not an ncnn .param file (magic 7767517 missing)" % param_path)
    data = np.fromfile(bin_path, dtype=np.uint8)
    pos = 0
    layers = []
    for ln in lines[2:]:
        t = ln.split()
        if not t or t[0] not in _CONV_TYPES:
            continue
        nin, nout = int(t[2]), int(t[3])
        kv = {int(k): v for k, v in (x.split("=") for x in t[4 + nin + nout:])}
        nw, no, has_bias = int(kv[6]), int(kv[0]), int(kv.get(5, 0)) != 0
        tag = int(np.frombuffer(data[pos:pos + 4].tobytes(), dtype=np.uint32)[0])
        if tag != 0:
            raise _lib.YfError("layer %s: weight storage tag 0x%x (fp16 / int8) is not supported, only raw fp32" % (t[1], tag))
        pos += 4
        w = np.frombuffer(data[pos:pos + 4 * nw].tobytes(), dtype=np.float32).copy()
        pos += 4 * nw
        b = None
        if has_bias:
            b = np.frombuffer(data[pos:pos + 4 * no].tobytes(), dtype=np.float32).copy()
            pos += 4 * no
        layers.append({"type": t[0], "name": t[1], "num_output": no, "kernel": int(kv.get(1, 1)), "stride": int(kv.get(3, 1)),
                       "group": int(kv.get(7, 1)), "weight": w, "bias": b})
    if pos != data.size:
        raise _lib.YfError("%s: %d bytes left after the last convolution (not the network of %s?)" % (bin_path, data.size - pos, param_path))
    return layers
