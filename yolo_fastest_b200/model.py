"""Drop-in for the reference model class (src/model_training/model/yolo_fastest.py:69-231).

``YoloFastest(io_params)`` keeps the reference's constructor, attribute names and state_dict keys
(508 entries, e.g. ``res2_1.conv2.0.weight``, ``conv0.1.running_var``, ``head_5.bias``), so
``load_state_dict(torch.load("models/pytorch/.../YOLO-Fastest_epoch_28.pth"))`` works unchanged
(detect.py:89-91).  The parameters are held in ordinary ``nn`` modules only as storage: ``forward``
folds BatchNorm on the host (fp64), hands the packed fp32 blob to libyf_b200.so once, and from then
on every call is ``yf_forward`` — hand-written sm_100a kernels, no PyTorch operator on the path.
There is no CPU or eager fallback: a CPU tensor, training mode or a missing library raises.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-5

# (attribute, kind, cin, cout, kernel, stride, depthwise, relu) in forward order (yolo_fastest.py:78-148).
# kind: "cbr" conv+BN(+ReLU) Sequential, "res" BasicResBlock(io, inner), "head" biased 1x1 conv, "up" deconv+BN+ReLU
ARCH = [
    ("conv0", "cbr", None, 8, 3, 2, False, True),
    ("conv1_2", "cbr", 8, 8, 1, 1, False, True),
    ("conv1_3", "cbr", 8, 8, 3, 1, True, True),
    ("conv1_4", "cbr", 8, 4, 1, 1, False, False),
    ("res1_1", "res", 4, 8, 0, 0, False, False),
    ("conv1_8", "cbr", 4, 24, 1, 1, False, True),
    ("conv1_9", "cbr", 24, 24, 3, 2, False, True),
    ("conv2_1", "cbr", 24, 8, 1, 1, False, False),
    ("res2_1", "res", 8, 32, 0, 0, False, False),
    ("res2_2", "res", 8, 32, 0, 0, False, False),
    ("conv2_2", "cbr", 8, 32, 1, 1, False, True),
    ("conv2_3", "cbr", 32, 32, 3, 2, True, True),
    ("conv3_1", "cbr", 32, 8, 1, 1, False, False),
    ("res3_1", "res", 8, 48, 0, 0, False, False),
    ("res3_2", "res", 8, 48, 0, 0, False, False),
    ("conv3_2", "cbr", 8, 48, 1, 1, False, True),
    ("conv3_3", "cbr", 48, 48, 3, 1, True, True),
    ("conv3_4", "cbr", 48, 16, 1, 1, False, False),
    ("res3_3", "res", 16, 96, 0, 0, False, False),
    ("res3_4", "res", 16, 96, 0, 0, False, False),
    ("res3_5", "res", 16, 96, 0, 0, False, False),
    ("res3_6", "res", 16, 96, 0, 0, False, False),
    ("conv3_5", "cbr", 16, 96, 1, 1, False, True),
    ("conv3_6", "cbr", 96, 96, 3, 2, True, True),
    ("conv4_1", "cbr", 96, 24, 1, 1, False, False),
    ("res4_1", "res", 24, 136, 0, 0, False, False),
    ("res4_2", "res", 24, 136, 0, 0, False, False),
    ("res4_3", "res", 24, 136, 0, 0, False, False),
    ("res4_4", "res", 24, 136, 0, 0, False, False),
    ("conv4_2", "cbr", 24, 136, 1, 1, False, True),
    ("conv4_3", "cbr", 136, 136, 3, 2, True, True),
    ("conv5_1", "cbr", 136, 48, 1, 1, False, True),
    ("res5_1", "res", 48, 224, 0, 0, False, False),
    ("res5_2", "res", 48, 224, 0, 0, False, False),
    ("res5_3", "res", 48, 224, 0, 0, False, False),
    ("res5_4", "res", 48, 224, 0, 0, False, False),
    ("res5_5", "res", 48, 224, 0, 0, False, False),
    ("conv5_2", "cbr", 48, 96, 1, 1, False, True),
    ("conv5_3", "cbr", 96, 96, 5, 1, True, True),
    ("conv5_4", "cbr", 96, 128, 1, 1, False, False),
    ("conv5_5", "cbr", 128, 128, 5, 1, True, True),
    ("conv5_6", "cbr", 128, 128, 1, 1, False, False),
    ("head_5", "head", 128, None, 1, 1, False, False),
    ("deconv5_1", "up", 96, 96, 2, 2, False, True),
    ("conv4_1_1", "cbr", 232, 96, 1, 1, False, True),
    ("conv4_1_2", "cbr", 96, 96, 5, 1, True, True),
    ("conv4_1_3", "cbr", 96, 96, 1, 1, False, False),
    ("conv4_1_4", "cbr", 96, 96, 5, 1, True, True),
    ("conv4_1_5", "cbr", 96, 96, 1, 1, False, False),
    ("head_4", "head", 96, None, 1, 1, False, False),
]


def _cbr(cin, cout, k, stride, depthwise, relu):
    layers = [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=cin if depthwise else 1, bias=False),
              nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU(inplace=False))
    return nn.Sequential(*layers)


class BasicResBlock(nn.Module):
    """Parameter container with the reference block's attribute names (yolo_fastest.py:52-58)."""

    def __init__(self, io_channels, inner_channels):
        super().__init__()
        self.conv1 = _cbr(io_channels, inner_channels, 1, 1, False, True)
        self.conv2 = _cbr(inner_channels, inner_channels, 3, 1, True, True)
        self.conv3 = _cbr(inner_channels, io_channels, 1, 1, False, False)

    def forward(self, x):
        raise _lib.YfError("BasicResBlock is a parameter container; the block runs fused inside yf_forward")


def _fold(conv_w, bn, transposed=False):
    """Eval-mode BN folded into the conv in fp64: w' = w*g/sqrt(v+eps), b' = beta - mean*g/sqrt(v+eps). The parameters live on the
    host (YoloFastest.to / .cuda do not move them), so nothing here touches the device."""
    w = conv_w.detach().double()
    g, b = bn.weight.detach().double(), bn.bias.detach().double()
    m, v = bn.running_mean.detach().double(), bn.running_var.detach().double()
    s = g / torch.sqrt(v + bn.eps)
    w = w * (s.view(1, -1, 1, 1) if transposed else s.view(-1, 1, 1, 1))
    return w.float().numpy().ravel(), (b - m * s).float().numpy().ravel()


class YoloFastest(nn.Module):
    VARIANT = _lib.VARIANT_FULL

    def __init__(self, io_params):
        super().__init__()
        self.num_cls = io_params["num_cls"]
        self.input_channel = io_params["input_channel"]
        self.num_anchors = io_params["num_anchors"]
        # yolo_fastest.py:74-75; the lite class multiplies its anchor count by the class count first (:240-241)
        self.num_out = (self.num_anchors * self.num_cls if self.VARIANT == _lib.VARIANT_LITE else self.num_anchors) * (5 + self.num_cls)
        for name, kind, cin, cout, k, s, dw, relu in ARCH:
            if kind == "cbr":
                mod = _cbr(self.input_channel if cin is None else cin, cout, k, s, dw, relu)
            elif kind == "res":
                mod = BasicResBlock(cin, cout)
            elif kind == "head":
                mod = nn.Conv2d(cin, self.num_out, kernel_size=1, stride=1)
            else:
                mod = nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=s, padding=0, bias=False),
                                    nn.BatchNorm2d(cout), nn.ReLU())
            setattr(self, name, mod)
        self._ctx = None          # _lib.Ctx for the current (device, H, W, max_batch)
        self._dirty = True        # packed weights on the device are stale

    # ---- parameter bookkeeping -------------------------------------------------------------
    def initialize_weights(self):
        """Same initialisation as the reference (yolo_fastest.py:220-231)."""
        for m in self.modules():
            if type(m) is nn.Conv2d:
                nn.init.kaiming_normal_(m.weight.data, nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif type(m) is nn.BatchNorm2d:
                m.weight.data.normal_(1.0, 0.02)
                m.bias.data.fill_(0)
        self._dirty = True

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        for m in self.modules():                      # load_ncnn turns the BatchNorms into identities with eps 0: a checkpoint
            if isinstance(m, nn.BatchNorm2d):         # brings real statistics back, and the reference's eps with them
                m.eps = 1e-5
        self._dirty = True
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._dirty = True
        return out

    # The 508 state_dict tensors are STORAGE for the host-side fold; the device only ever sees the packed blob (yf_load_weights).
    # `.to(device)` / `.cuda()` therefore keep them on the host (detect.py:89 calls `.to(device)`; moving 508 small tensors to the GPU
    # and pulling each back to fold it cost hundreds of tiny copies and syncs per context). dtype conversions still apply.
    def to(self, *args, **kwargs):
        device, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if dtype is not None:
            super().to(dtype=dtype)
        return self

    def cuda(self, device=None):
        return self

    def refresh(self):
        """Call after mutating parameters in place; the next forward re-folds and re-uploads them."""
        self._dirty = True

    def folded_blob(self):
        """BN-folded fp32 parameters in the order yf_load_weights expects (include/yf.h)."""
        parts = []
        for name, kind, *_ in ARCH:
            mod = getattr(self, name)
            if kind == "cbr":
                parts.extend(_fold(mod[0].weight, mod[1]))
            elif kind == "res":
                for sub in (mod.conv1, mod.conv2, mod.conv3):
                    parts.extend(_fold(sub[0].weight, sub[1]))
            elif kind == "head":
                parts.append(mod.weight.detach().float().cpu().numpy().ravel())
                parts.append(mod.bias.detach().float().cpu().numpy().ravel())
            else:
                parts.extend(_fold(mod[0].weight, mod[1], transposed=True))
        blob = np.concatenate(parts).astype(np.float32)
        expect = _lib.lib().yf_weight_count_variant(self.input_channel, self.num_cls, self.num_anchors, self.VARIANT)
        if blob.size != expect:
            raise _lib.YfError("folded blob has %d floats, library expects %d" % (blob.size, expect))
        return blob

    def load_ncnn(self, param_path, bin_path):
        """Take the weights from the reference's ncnn deployment files (`models/ncnn/<res>/*-opt.param|bin`) instead
        of a `.pth`. The ncnn network is already BN-folded, so every conv receives the folded kernel and its BatchNorm becomes the
        identity carrying the folded bias (weight 1, running_mean 0, running_var 1, eps 0, bias b'): `folded_blob` then reproduces
        the ncnn arrays bit for bit. The layer table is checked against the architecture first."""
        from . import ncnn_loader
        layers = ncnn_loader.read_ncnn(param_path, bin_path)
        slots = []                                   # (conv module, bn module or None, transposed)
        for name, kind, *_ in ARCH:
            mod = getattr(self, name)
            if kind == "cbr":
                slots.append((mod[0], mod[1], False))
            elif kind == "res":
                slots.extend((sub[0], sub[1], False) for sub in (mod.conv1, mod.conv2, mod.conv3))
            elif kind == "head":
                slots.append((mod, None, False))
            else:
                slots.append((mod[0], mod[1], True))
        if len(layers) != len(slots):
            raise _lib.YfError("ncnn file has %d convolutions, YOLO-Fastest has %d" % (len(layers), len(slots)))
        with torch.no_grad():
            for (conv, bn, transposed), l in zip(slots, layers):
                w = torch.from_numpy(l["weight"])
                cout = conv.weight.shape[1] * conv.groups if transposed else conv.weight.shape[0]
                if w.numel() != conv.weight.numel() or l["num_output"] != cout or l["kernel"] != conv.kernel_size[0] \
                        or l["bias"] is None or transposed != l["type"].startswith("Deconv"):
                    raise _lib.YfError("ncnn layer %s (%s, %d outputs, kernel %d, %d weights) does not match %s"
                                       % (l["name"], l["type"], l["num_output"], l["kernel"], w.numel(), tuple(conv.weight.shape)))
                if transposed:                       # ncnn stores [out][in][kh][kw], ConvTranspose2d [in][out][kh][kw]
                    ci, co, kh, kw = conv.weight.shape
                    w = w.view(co, ci, kh, kw).permute(1, 0, 2, 3).contiguous()
                conv.weight.copy_(w.view_as(conv.weight))
                b = torch.from_numpy(l["bias"])
                if bn is None:
                    conv.bias.copy_(b)
                else:
                    bn.weight.fill_(1.0)
                    bn.bias.copy_(b)
                    bn.running_mean.zero_()
                    bn.running_var.fill_(1.0)
                    bn.eps = 0.0
        self._dirty = True
        return self

    def _slots(self):
        """(conv module, bn module or None, transposed) per convolution, in forward order"""
        slots = []
        for name, kind, *_ in ARCH:
            mod = getattr(self, name)
            if kind == "cbr":
                slots.append((mod[0], mod[1], False))
            elif kind == "res":
                slots.extend((sub[0], sub[1], False) for sub in (mod.conv1, mod.conv2, mod.conv3))
            elif kind == "head":
                slots.append((mod, None, False))
            else:
                slots.append((mod[0], mod[1], True))
        return slots

    def load_onnx(self, path):
        """Take the parameters from the reference's ONNX export (`models/onnx/<res>/*.onnx`): convolution kernels,
        BatchNorm scale / bias / mean / var / epsilon and the head biases go into the module's own tensors, so `folded_blob` folds them
        exactly as it folds a `.pth`. The node table is checked against the architecture first."""
        from . import onnx_loader
        layers = onnx_loader.read_onnx(path)
        slots = self._slots()
        if len(layers) != len(slots):
            raise _lib.YfError("ONNX file has %d convolutions, YOLO-Fastest has %d" % (len(layers), len(slots)))
        with torch.no_grad():
            for (conv, bn, transposed), l in zip(slots, layers):
                w = torch.from_numpy(l["weight"])
                if tuple(w.shape) != tuple(conv.weight.shape) or transposed != (l["type"] == "ConvTranspose") or (bn is None) != (l["bn"] is None) \
                        or l["group"] != conv.groups or l["stride"] != conv.stride[0]:
                    raise _lib.YfError("ONNX node %s (%s %s, stride %d, group %d) does not match %s"
                                       % (l["name"], l["type"], tuple(w.shape), l["stride"], l["group"], tuple(conv.weight.shape)))
                conv.weight.copy_(w)
                if bn is None:
                    conv.bias.copy_(torch.from_numpy(l["bias"]))
                else:
                    bn.weight.copy_(torch.from_numpy(l["bn"]["scale"]))
                    bn.bias.copy_(torch.from_numpy(l["bn"]["bias"]))
                    bn.running_mean.copy_(torch.from_numpy(l["bn"]["mean"]))
                    bn.running_var.copy_(torch.from_numpy(l["bn"]["var"]))
                    bn.eps = l["bn"]["eps"]
        self._dirty = True
        return self

    # ---- execution ---------------------------------------------------------------------------
    def context(self, device, H, W, batch):
        """The yf_ctx serving (device, H, W); grown (re-created) when a larger batch arrives."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        c = self._ctx
        if c is None or c.device_index != idx or c.H != H or c.W != W or c.max_batch < batch:
            if c is not None:
                if c.pending:
                    raise _lib.YfError("a larger batch needs a new context, but slots %s hold submitted batches: collect them first" % sorted(c.pending))
                torch.cuda.synchronize(c.device_index)
                c.close()
            with torch.cuda.device(idx):
                c = _lib.Ctx(idx, self.input_channel, self.num_cls, self.num_anchors, batch, H, W, self.VARIANT)
            self._ctx = c
            self._dirty = True
        if self._dirty:
            with torch.cuda.device(idx):
                torch.cuda.synchronize(idx)
                c.load_weights(self.folded_blob())
            self._dirty = False
        return c

    def _check_input(self, x):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise _lib.YfError("YoloFastest.forward runs only on CUDA tensors (no CPU fallback); got %s"
                               % (x.device if isinstance(x, torch.Tensor) else type(x)))
        if self.training:
            raise _lib.YfError("the B200 path is inference-only (BatchNorm is folded): call .eval() first")
        if x.dim() != 4 or x.shape[1] != self.input_channel:
            raise _lib.YfError("expected input [B, %d, H, W], got %s" % (self.input_channel, tuple(x.shape)))
        if x.shape[2] % 32 or x.shape[3] % 32:
            raise _lib.YfError("H and W must be multiples of 32 (_config.py:11), got %dx%d" % (x.shape[2], x.shape[3]))

    def forward(self, x):
        """[B, C, H, W] fp32 cuda -> (head_large [B, A(5+nc), H/16, W/16], head_small [.., H/32, W/32])."""
        self._check_input(x)
        x = x.contiguous().float()
        B, _, H, W = x.shape
        ctx = self.context(x.device, H, W, B)
        head_large = torch.empty((B, self.num_out, H // 16, W // 16), dtype=torch.float32, device=x.device)
        head_small = torch.empty((B, self.num_out, H // 32, W // 32), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().yf_forward(ctx.handle, x.data_ptr(), B, head_large.data_ptr(), head_small.data_ptr(),
                                             C.c_void_p(stream)), ctx.handle)
        return head_large, head_small

    def tap(self, name, batch):
        """Intermediate activation ``name`` of the last forward (parity/debug aid; see yf_tap)."""
        ctx = self._ctx
        per = C.c_int64(0)
        _lib.check(_lib.lib().yf_tap(ctx.handle, name.encode(), batch, None, C.byref(per), None), ctx.handle)
        out = torch.empty((batch, per.value), dtype=torch.float32, device="cuda:%d" % ctx.device_index)
        stream = torch.cuda.current_stream(out.device).cuda_stream
        _lib.check(_lib.lib().yf_tap(ctx.handle, name.encode(), batch, out.data_ptr(), C.byref(per), C.c_void_p(stream)), ctx.handle)
        return out

    def profile(self, x):
        """Per-group device milliseconds of one forward (yf_profile_forward)."""
        self._check_input(x)
        x = x.contiguous().float()
        B, _, H, W = x.shape
        ctx = self.context(x.device, H, W, B)
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        torch.cuda.synchronize(x.device)
        with torch.cuda.device(x.device):
            n = _lib.lib().yf_profile_forward(ctx.handle, x.data_ptr(), B, names, ms, 64)
        if n < 0:
            _lib.check(n, ctx.handle)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]


class YoloFastest_lite(YoloFastest):
    """Drop-in for the reference's single-head variant (yolo_fastest.py:234-387): the same parameter set and state_dict keys as
    YoloFastest with (num_anchors * num_cls) * (5 + num_cls) head channels (:240-241); its forward goes conv3_2 -> conv3_4 without the
    depthwise conv3_3 (:335-337), ends at head_5 and returns that single tensor (:365-372). Same kernels, one more launch plan
    (YF_VARIANT_LITE); the reference's `initialize_weights` additionally sets BatchNorm eps 1e-3 / momentum 0.03 (:383-385)."""
    VARIANT = _lib.VARIANT_LITE

    def initialize_weights(self):
        super().initialize_weights()
        for m in self.modules():
            if type(m) is nn.BatchNorm2d:
                m.eps = 1e-3
                m.momentum = 0.03
        self._dirty = True

    def load_state_dict(self, *args, **kwargs):
        eps = {k: m.eps for k, m in self.named_modules() if isinstance(m, nn.BatchNorm2d)}
        out = super().load_state_dict(*args, **kwargs)
        for k, m in self.named_modules():              # a checkpoint does not carry eps: keep what the instance had
            if isinstance(m, nn.BatchNorm2d):
                m.eps = eps[k]
        return out

    def forward(self, x):
        """[B, 1, H, W] fp32 cuda -> head_5 [B, A*nc*(5+nc), H/32, W/32]."""
        self._check_input(x)
        x = x.contiguous().float()
        B, _, H, W = x.shape
        ctx = self.context(x.device, H, W, B)
        head = torch.empty((B, self.num_out, H // 32, W // 32), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().yf_forward(ctx.handle, x.data_ptr(), B, None, head.data_ptr(), C.c_void_p(stream)), ctx.handle)
        return head
