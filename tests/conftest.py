import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden_cases(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b, floor=1e-12):
    """max |a-b| / max(|b|, floor): the parity metric of BASELINE.md §3.6 (abs floor 1e-12)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape
    if a.size == 0:
        return 0.0
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), floor))) if m.any() else 0.0


@pytest.fixture(scope="session")
def oracle():
    from oracle import lisf_oracle
    lisf_oracle.lib()
    return lisf_oracle


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, initialised on device 0.  Fails (does not skip) when it cannot run:
    -m gpu tests must never pass on a fallback."""
    from lisflood_code_b200 import _capi
    L = _capi.lib()
    _capi.check(L.lf_device_init(0))
    return _capi


def golden_model(name):
    """(S, [F per step], [outputs per step]) of a model_* golden case."""
    g = load_golden(name)
    S, steps = {}, int(g["steps"])
    for k, v in g.items():
        if k.startswith("S__"):
            S[k[3:]] = v.item() if v.ndim == 0 else v
    S["N"], S["rows"], S["cols"], S["NoRoutSteps"] = int(S["N"]), int(S["rows"]), int(S["cols"]), int(S["NoRoutSteps"])
    S["SplitRouting"] = bool(S["SplitRouting"])
    F = [{k.split("__", 1)[1]: v for k, v in g.items() if k.startswith("F%d__" % t)} for t in range(steps)]
    O = [{k.split("__", 1)[1]: v for k, v in g.items() if k.startswith("O%d__" % t)} for t in range(steps)]
    return S, F, O
