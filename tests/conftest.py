import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C-ABI library must exist for every test session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def gold():
    class G:
        dir = GOLD
        res = {r: np.load(os.path.join(GOLD, "golden_%s.npz" % r)) for r in ("256x320", "512x640")}
        stress = np.load(os.path.join(GOLD, "golden_stress80_416.npz"))

        @staticmethod
        def ckpt(name):
            return os.path.join(GOLD, "weights", name + ".pth")

        @staticmethod
        def sd(name):
            return torch.load(os.path.join(GOLD, "weights", name + ".pth"), map_location="cpu")
    return G


def rows_equal(want, got, conf_tol=1e-9):
    """Compare detection rows [x1,y1,x2,y2,conf,cls_score,cls(,src)]: integer fields exact, floats within conf_tol."""
    assert len(want) == len(got), "row count %d != %d" % (len(want), len(got))
    for i, (w, g) in enumerate(zip(want, got)):
        assert [int(v) for v in w[:4]] == [int(v) for v in g[:4]], "row %d box %s != %s" % (i, list(w[:4]), list(g[:4]))
        assert int(w[6]) == int(g[6]), "row %d class" % i
        assert abs(float(w[4]) - float(g[4])) <= conf_tol * max(1.0, abs(float(w[4]))), "row %d conf %r != %r" % (i, w[4], g[4])
        assert abs(float(w[5]) - float(g[5])) <= conf_tol * max(1.0, abs(float(w[5]))), "row %d cls_score" % i
