"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/*.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        names |= set(re.findall(r"ADB_API[^;(]*?\b(adb_\w+)\s*\(", src))
    return sorted(names)


@pytest.fixture(scope="module")
def lib():
    import analytical_database_b200 as adb
    if not os.path.exists(adb.lib_path()):
        adb.build_native()
    return C.CDLL(adb.lib_path())


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from analytical_database_b200.engine import Engine
    assert sorted(Engine._SIGS) == declared_symbols()


def test_no_cpu_fallback(lib):
    """Without adb_init() (or without a GPU) operators fail loudly instead of computing."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    lib.adb_last_error.restype = C.c_char_p
    lib.adb_sync.restype = C.c_int32
    assert lib.adb_sync() == -1                      # ADB_ERR_NOT_INITIALISED
    assert b"adb_init" in lib.adb_last_error()
    if not has_gpu:
        lib.adb_init.restype = C.c_int32
        assert lib.adb_init(0) == -2                 # ADB_ERR_CUDA: no device
        assert b"no CPU fallback" in lib.adb_last_error()
        import analytical_database_b200 as adb
        with pytest.raises(adb.EngineError):
            adb.Engine(0)


def test_product_never_touches_the_oracle():
    """No file of the engine package references oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "analytical-database_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in text and "import oracle" not in text \
                    and "from oracle" not in text and "orc_" not in text, os.path.join(base, f)
