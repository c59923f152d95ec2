"""B200-native operator engine for the reference column store's data-parallel path.

Layout (only what the hot path needs):

* ``csrc/``       hand-written sm_100a kernels + the C-ABI (``include/adb_engine.h``),
                  built in-tree into ``libadb_b200.so``;
* ``host/``       the C host side: ``query_shim.c`` implements the reference's
                  ``query.h`` operator API on top of the C-ABI (the drop-in);
* ``engine.py``   ctypes binding of the C-ABI for tests / bench (plumbing only);
* ``sharded.py``  row-range sharding over N GPUs: count / aggregate / join exchange steps
                  on ``torch.distributed`` (one process per GPU);
* ``synth.py``    the counter-based synthetic column generator (numpy twin of
                  ``adb_synth_uniform``).

There is no CPU fallback: importing works anywhere, but ``Engine()`` raises unless the
CUDA library is built and a device is present.
"""
from .engine import Engine, EngineError, DevBuf, Agg, lib_path, build_native  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["Engine", "EngineError", "DevBuf", "Agg", "lib_path", "build_native", "synth"]
