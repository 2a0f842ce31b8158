import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tolerances stated by BASELINE.json north_star (normwise relative error vs the float64 reference)
TOL = {"fp32": 1e-5, "tf32": 5e-3, "bf16": 2e-2}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """Return (arrays, cases) of a fixture written by tests/golden/make_golden.py."""
    z = np.load(os.path.join(GOLDEN, name))
    cases = json.loads(str(z["__cases__"]))
    return z, cases


def to_pad(p):
    return tuple(p) if isinstance(p, list) else p


def rel_err(a, ref):
    """Normwise relative error ||a - ref||_2 / ||ref||_2 (absolute if ref == 0)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    den = np.linalg.norm(ref.ravel())
    num = np.linalg.norm((a - ref).ravel())
    return num / den if den > 0 else num


def npml():
    """The product package (its directory name is not an identifier)."""
    return importlib.import_module("numpy-ml_b200")


@pytest.fixture(scope="session")
def pkg():
    return npml()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
