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


# Measured parity figures that are REPORTED, not hidden behind a tolerance (violation counts of the survey's strict allclose form,
# integer-box agreement counts, distances from float64): tests call report(key, value); the session prints them in the terminal
# summary and writes gpurun_out/parity_report.json (copied to profiles/ by the builder).
PARITY_REPORT = {}


def report(key, value):
    PARITY_REPORT[key] = value


def pytest_terminal_summary(terminalreporter):
    if not PARITY_REPORT:
        return
    import json
    terminalreporter.write_line("parity report (measured, see tests/conftest.py):")
    for k in sorted(PARITY_REPORT):
        terminalreporter.write_line("  %s: %s" % (k, PARITY_REPORT[k]))
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        json.dump(PARITY_REPORT, open(os.path.join(out, "parity_report.json"), "w"), indent=1, sort_keys=True)
    except OSError:
        pass


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
