"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on a CPU-only host: oracle vs reference vs golden vectors, host
logic, C-ABI symbol checks.  `-m gpu` are the parity tests proper: they call the CUDA
engine through its C-ABI and compare against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle
    r = oracle.reference("O2")
    if r is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return r


@pytest.fixture(scope="session")
def ref_o0():
    from oracle import oracle
    r = oracle.reference("O0")
    if r is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return r


@pytest.fixture
def rng():
    return np.random.default_rng(42)
