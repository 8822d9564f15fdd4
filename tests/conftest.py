import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "exact_gp_sklearn.npz")


@pytest.fixture(scope="session")
def pre_golden():
    """Outputs of the reference's own PreProcessor (tests/golden/make_golden_reference.py)."""
    return np.load(ROOT / "tests" / "golden" / "preprocess_reference.npz")


@pytest.fixture(scope="session")
def met_golden():
    """Outputs of the reference's own gpras/metrics.py (tests/golden/make_golden_reference.py)."""
    return np.load(ROOT / "tests" / "golden" / "metrics_reference.npz")


PRE_CASES = ["wse_fixed", "depth_fixed", "wse_north", "velocity"]
MET_CASES = ["small", "ragged", "one_step", "peaks_differ"]


def sub(npz, name):
    """All arrays of one golden case as a dict."""
    pre = name + "."
    return {k[len(pre):]: npz[k] for k in npz.files if k.startswith(pre)}


GOLDEN_CASES = ["rbf_iso", "rbf_ard", "m12_iso", "m32_ard", "m52_ard", "m52_iso_dup"]


def golden_case(golden, name):
    keys = ["x", "y", "xs", "variance", "noise", "ls", "lml", "grad_log", "mean", "std"]
    c = {k: golden[f"{name}.{k}"] for k in keys}
    c["kernel"] = str(golden[f"{name}.kernel"])
    c["variance"], c["noise"], c["lml"] = float(c["variance"]), float(c["noise"]), float(c["lml"])
    return c


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    from gpras_b200 import _lib, build

    build.build()
    return _lib.load()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must never silently pass on a box without a GPU
    pass
