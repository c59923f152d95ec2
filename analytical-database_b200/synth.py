"""Counter-based synthetic columns: the numpy twin of ``adb_synth_uniform``
(csrc/gather_agg.cu, ``mix64`` in csrc/adb_common.cuh).

``uniform(n, seed, first_row, lo, span)[i] == lo + ((mix64(seed, first_row+i) >> 32) * span >> 32)``
so any row range of any shard can be regenerated anywhere (GPU, host, another rank)
without materialising the whole column -- how bench.py and the tests give the CPU
oracle the same rows the GPU scanned.
"""
import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def mix64(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def uniform(n: int, seed: int, first_row: int = 0, lo: int = 0, span: int = 1 << 31,
            chunk: int = 1 << 24) -> np.ndarray:
    """int32 column of n rows uniform in [lo, lo + span)."""
    out = np.empty(n, dtype=np.int32)
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        z = mix64(seed, np.arange(first_row + b, first_row + e, dtype=np.uint64))
        v = ((z >> np.uint64(32)) * np.uint64(span)) >> np.uint64(32)
        out[b:e] = (v.astype(np.int64) + lo).astype(np.int32)
    return out
