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


class PinnedOracle:
    """The restatement (oracle/adb_oracle.c) with the unmodified reference objects (oracle/_ref)
    run beside it: for the operators the reference computes without reading out of bounds,
    every call is made on BOTH and the answers must be identical before the restatement's is
    returned.  GPU parity tests that take `port` are thereby checked against the reference
    itself wherever oracle/_ref exists (this container; on the GPU box the prebuilt objects
    travel with the snapshot)."""
    SAFE = ("select_scan", "select_result", "fetch", "sum", "add", "sub", "chain_select_fetch_sum")

    def __init__(self, port, ref):
        self._port, self._ref = port, ref

    def __getattr__(self, name):
        fn = getattr(self._port, name)
        if self._ref is None or name not in self.SAFE:
            return fn
        rfn = getattr(self._ref, name)

        def both(*a, **kw):
            x, y = fn(*a, **kw), rfn(*a, **kw)
            same = np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y
            assert same, f"oracle restatement and reference disagree on {name}"
            return x
        return both


@pytest.fixture(scope="session")
def port():
    from oracle import oracle
    return PinnedOracle(oracle.port(), oracle.reference("O2"))


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
