import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")

# parity tolerances (BASELINE.json north_star): samples within 1e-10 relative of the reference;
# constraint residual ||Gx - r|| within 1e-12 relative (to ||r||, or absolute when r = 0)
REL_TOL = 1e-10
RESID_TOL = 1e-12


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(x, ref):
    """norm-wise relative error per sample row: max|x - ref| / max|ref|"""
    x, ref = np.asarray(x), np.asarray(ref)
    scale = np.maximum(np.abs(ref).max(axis=-1, keepdims=True), 1e-300)
    return float((np.abs(x - ref) / scale).max())


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import Reference, have_reference
    if not have_reference() and not os.path.isdir("/root/reference"):
        pytest.skip("compiled reference (oracle/_ref) not available")
    return Reference()


@pytest.fixture(scope="session")
def gold():
    return {name: np.load(os.path.join(GOLD, name + ".npz")) for name in ("rng", "hyperplane", "structured")}
