"""Result text on the device: adb_format_i32_count / _emit against the
C library's "%d" -- the INT branch of print, /root/reference/src/query.c:262-269: one "%d"
per tuple, "\\n" between tuples, nothing after the last."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
I32MIN, I32MAX = -2**31, 2**31 - 1


@pytest.fixture(scope="module")
def eng():
    from analytical_database_b200 import Engine
    return Engine(0)


def expected(vals) -> bytes:
    return "\n".join("%d" % int(v) for v in vals).encode()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 31, 255, 1023, 1024, 1025, 4096, 100_003, 1_048_576])
def test_format_matches_printf(eng, rng, n):
    # every width from 1 to 10 digits, both signs
    mag = rng.integers(0, 11, n)
    vals = (rng.integers(0, 10, n) * 10 ** np.minimum(mag, 9) + rng.integers(0, 1000, n)) * rng.choice([-1, 1], n)
    vals = np.clip(vals, I32MIN, I32MAX).astype(np.int32)
    d = eng.upload(vals)
    assert eng.format_i32(d, n) == expected(vals)
    d.free()


def test_format_equals_oracle_print(eng, port, rng):
    for n in (1, 5000, 70_001):
        vals = rng.integers(I32MIN, I32MAX, n).astype(np.int32)
        d = eng.upload(vals)
        assert eng.format_i32(d, n) == port.print_i32(vals)
        d.free()


def test_format_extremes_and_unaligned_views(eng):
    vals = np.array([0, -1, 1, 9, 10, -10, 99, 100, I32MAX, I32MIN, I32MIN + 1, 1000000000, -1000000000,
                     999999999, -999999999, 7] * 70, dtype=np.int32)
    d = eng.upload(vals)
    assert eng.format_i32(d, vals.size) == expected(vals)
    for off in (1, 2, 3, 5):                                  # value pointer not 16-byte aligned
        nb = C.c_int64(0)
        n = vals.size - off
        eng._ck(eng.lib.adb_format_i32_count(d.i32(off), n, C.byref(nb)))
        t = eng.alloc(nb.value)
        eng._ck(eng.lib.adb_format_i32_emit(t.void()))
        assert t.to_host(nb.value, np.uint8).tobytes() == expected(vals[off:])
        t.free()
    d.free()


def test_format_protocol_errors(eng):
    from analytical_database_b200 import EngineError
    with pytest.raises(EngineError, match="no preceding"):
        eng._ck(eng.lib.adb_format_i32_emit(None))
    d = eng.upload(np.arange(10, dtype=np.int32))
    nb = C.c_int64(0)
    eng._ck(eng.lib.adb_format_i32_count(d.i32(), 10, C.byref(nb)))
    assert nb.value == 19
    t = eng.alloc(64)
    with pytest.raises(EngineError, match="aligned"):
        eng._ck(eng.lib.adb_format_i32_emit(t.void(4)))
    d.free()
    t.free()


def test_format_large_result_round_trip(eng):
    """20 M values (~200 MB of text through the staged download): parse it back."""
    n = 20_000_000
    d = eng.synth_uniform(n, 5, 0, -2**30, 2**31 - 1)
    text = eng.format_i32(d, n)
    back = np.fromstring(text.decode(), sep="\n", dtype=np.int64)
    assert np.array_equal(back.astype(np.int32), d.to_host(n))
    d.free()
