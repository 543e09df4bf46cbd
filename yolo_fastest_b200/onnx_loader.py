"""Weights from the reference's ONNX exports: `models/onnx/<res>/*.onnx`, written by
`src/model_deployment/convert_model/pytorch_to_onnx.py` from the same checkpoints. The graph keeps the BatchNorms as nodes
(85 Conv + 1 ConvTranspose, 84 BatchNormalization, 57 Relu, 18 Add, 1 Concat): every bias-free `Conv` / `ConvTranspose` is followed
by the `BatchNormalization` that consumes its output (scale, B, mean, var initializers + `epsilon`), the two head convs carry a
bias instead. `YoloFastest.load_onnx` puts these tensors into the module's convs and BatchNorms, so `folded_blob` folds them exactly
as it folds a `.pth` — the ONNX file is an independent serialisation of the same parameters, in forward order.

The `onnx` Python package is not a dependency: an ONNX file is a protobuf `ModelProto`, and the four message types needed here
(ModelProto.graph = 7; GraphProto.node = 1, .initializer = 5; NodeProto.input = 1, .output = 2, .name = 3, .op_type = 4,
.attribute = 5; TensorProto.dims = 1, .data_type = 2, .float_data = 4, .name = 8, .raw_data = 9; AttributeProto.name = 1, .f = 2,
.i = 3, .ints = 8) are read with a 40-line wire-format decoder."""
import struct

import numpy as np

from . import _lib


def _varint(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _fields(buf):
    """yield (field number, wire type, value) of one message: varint -> int, 64/32-bit -> bytes, length-delimited -> memoryview"""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 5:
            v, pos = bytes(buf[pos:pos + 4]), pos + 4
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        else:
            raise _lib.YfError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _packed_ints(v, wt):
    if wt == 0:
        return [v]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _tensor(buf):
    dims, dtype, name, raw, floats = [], 0, "", None, []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_ints(v, wt)
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats += list(struct.unpack("<%df" % (len(v) // 4), bytes(v))) if wt == 2 else [struct.unpack("<f", v)[0]]
    if dtype != 1:
        return name, None                      # only fp32 tensors matter here
    a = np.frombuffer(raw, dtype="<f4").copy() if raw is not None else np.asarray(floats, dtype=np.float32)
    return name, a.reshape(dims) if dims else a


def _node(buf):
    n = {"input": [], "output": [], "name": "", "op": "", "attr": {}}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n["input"].append(bytes(v).decode())
        elif fno == 2:
            n["output"].append(bytes(v).decode())
        elif fno == 3:
            n["name"] = bytes(v).decode()
        elif fno == 4:
            n["op"] = bytes(v).decode()
        elif fno == 5:
            an, ai, ints = "", None, []
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    an = bytes(v2).decode()
                elif f2 == 2 and w2 == 5:
                    ai = struct.unpack("<f", v2)[0]
                elif f2 == 3:
                    ai = v2
                elif f2 == 8:
                    ints += _packed_ints(v2, w2)
            n["attr"][an] = ints if ints else ai
    return n


def read_onnx(path):
    """-> list of dicts {type ("Conv" | "ConvTranspose"), name, num_output, kernel, stride, group, weight (fp32 array in the ONNX =
    PyTorch layout: Conv [out][in/group][kh][kw], ConvTranspose [in][out][kh][kw]), bias (array or None), bn (None or dict scale,
    bias, mean, var, eps)} in graph (= forward) order"""
    data = memoryview(open(path, "rb").read())
    graph = None
    try:
        for fno, wt, v in _fields(data):
            if fno == 7 and wt == 2:
                graph = v
        inits, nodes = {}, []
        for fno, wt, v in _fields(graph if graph is not None else b""):
            if fno == 5:
                name, a = _tensor(v)
                if a is not None:
                    inits[name] = a
            elif fno == 1:
                nodes.append(_node(v))
    except (IndexError, struct.error, ValueError, _lib.YfError):
        graph = None
    if graph is None:
        raise _lib.YfError("%s is not an ONNX model (no readable graph)" % path)
    consumer = {}
    for n in nodes:
        for i in n["input"][:1]:
            consumer.setdefault(i, n)
    layers = []
    for n in nodes:
        if n["op"] not in ("Conv", "ConvTranspose"):
            continue
        if len(n["input"]) < 2 or n["input"][1] not in inits:
            raise _lib.YfError("node %s (%s) has no constant weight" % (n["name"], n["op"]))
        w = inits[n["input"][1]]
        b = inits.get(n["input"][2]) if len(n["input"]) > 2 else None
        nxt = consumer.get(n["output"][0])
        bn = None
        if nxt is not None and nxt["op"] == "BatchNormalization":
            sc, bb, mu, var = (inits[k] for k in nxt["input"][1:5])
            bn = {"scale": sc, "bias": bb, "mean": mu, "var": var, "eps": float(nxt["attr"].get("epsilon") or 1e-5)}
        if (b is None) == (bn is None):
            raise _lib.YfError("node %s: expected either a bias or a following BatchNormalization" % n["name"])
        cout = int(w.shape[1]) * int(n["attr"].get("group") or 1) if n["op"] == "ConvTranspose" else int(w.shape[0])
        layers.append({"type": n["op"], "name": n["name"] or n["output"][0], "num_output": cout, "kernel": int(w.shape[-1]),
                       "stride": int((n["attr"].get("strides") or [1])[0]), "group": int(n["attr"].get("group") or 1),
                       "weight": w, "bias": b, "bn": bn})
    if not layers:
        raise _lib.YfError("%s holds no convolution" % path)
    return layers
