"""The C-ABI library: loads, exports every symbol include/yf.h declares, sizes agree with the host packer,
and fails loudly (no fallback) where there is no GPU.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

import yolo_fastest_b200 as yf
from yolo_fastest_b200 import _lib

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "yf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 19 and "yf_forward" in names and "yf_postprocess" in names and "yf_detect_host_u8" in names
    l = C.CDLL(yf.LIB_PATH)
    for n in names:
        assert hasattr(l, n), "libyf_b200.so does not export %s" % n
    bound = {s[0] for s in _lib.SYMBOLS}
    assert bound == set(names), "ctypes table and header disagree: %s" % (bound ^ set(names))


def test_abi_version_and_struct_sizes():
    assert yf.lib().yf_abi_version() == 1
    assert C.sizeof(_lib.YfDet) == 56 and _lib.DET_DTYPE.itemsize == 56
    assert C.sizeof(_lib.YfPostParams) == 2 * 8 * 2 * 8 + 16 + 16


def test_weight_count_matches_packer():
    for nc in (3, 80):
        m = yf.YoloFastest({"num_cls": nc, "input_channel": 1, "num_anchors": 3})
        n_params = sum(p.numel() for p in m.parameters())
        blob = m.folded_blob()
        assert blob.size == yf.lib().yf_weight_count(1, nc, 3)
        # folding removes gamma (one of the two BN vectors) per normalised conv: 84 BN layers
        bn = sum(mod.weight.numel() for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d))
        assert blob.size == n_params - bn
    assert yf.lib().yf_weight_count(0, 3, 3) < 0
    assert yf.lib().yf_weight_count(3, 3, 3) == yf.lib().yf_weight_count(1, 3, 3) + 2 * 72      # conv0 [8][3][3][3] vs [8][1][3][3]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    h = C.c_void_p()
    rc = yf.lib().yf_create(C.byref(h), 0, 1, 3, 3, 1, 256, 320)
    assert rc == -2 and not h.value                                    # YF_ERR_CUDA, never a CPU fallback
    assert b"no CPU fallback" in yf.lib().yf_last_error(None)
    m = yf.YoloFastest({"num_cls": 3, "input_channel": 1, "num_anchors": 3}).eval()
    with pytest.raises(yf.YfError):
        m(torch.zeros(1, 1, 256, 320))
    pp = yf.YOLO_post_process(0.5, 0.2, 3, 3, yf.config_for("256x320")["io_params"]["anchors"], [256, 320, 1])
    with pytest.raises(yf.YfError):
        pp.decode_box((torch.zeros(1, 24, 16, 20), torch.zeros(1, 24, 8, 10)))
    with pytest.raises(yf.YfError):
        pp.non_maxium_supression([[0, 0, 1, 1, 0.9, 0.9, 0]])


def test_create_rejects_bad_arguments():
    h = C.c_void_p()
    l = yf.lib()
    assert l.yf_create(C.byref(h), 0, 2, 3, 3, 1, 256, 320) == -1       # in_ch is 1 or 3
    assert b"in_ch" in l.yf_last_error(None)
    assert l.yf_create(C.byref(h), 0, 1, 3, 3, 1, 250, 320) == -1       # not a multiple of 32
    assert l.yf_create(C.byref(h), 0, 1, 3, 9, 1, 256, 320) == -1       # too many anchors
    assert l.yf_create(None, 0, 1, 3, 3, 1, 256, 320) == -1
    assert l.yf_forward(None, None, 1, None, None, None) == -1
    assert l.yf_launch_count(None) == -1


def test_build_is_content_addressed_and_idempotent():
    """build() decides by a hash of the sources recorded beside each library, not by mtimes (a copy of the tree to the GPU box does not
    keep them in order — N ranks rebuilding the same .so at once was the failure mode), holds a file lock while it compiles and renames
    the result into place."""
    import os
    import __graft_entry__ as g
    g.build()
    sources = [os.path.join(g.CSRC, f) for f in sorted(os.listdir(g.CSRC)) if f.endswith((".cu", ".cuh"))]
    sources.append(os.path.join(g.ROOT, "include", "yf.h"))
    assert not g._stale(g.LIB, sources, g.NVCC_FLAGS)
    before = os.path.getmtime(g.LIB)
    os.utime(sources[0], None)                       # a newer mtime alone must not trigger a rebuild
    g.build()
    assert os.path.getmtime(g.LIB) == before
    assert g._stale(g.LIB, sources, g.NVCC_FLAGS + ["-DSOMETHING_ELSE"])
