"""Host logic of the batched shared scan (replaces shared_select, query.c:450-583), checked on
the CPU: the lookup tables adb_shared_select_count uploads -- bounds, cover lists, value ->
interval table, prefilter bitmap, per-query interval ranges, the coloured cover table of the
pair lists -- classify every value exactly as the reference's predicate `low <= v < high`
(query.c:474) does.  No device, no compute call: adb_shared_select_plan is pure host code."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I32 = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def lib():
    import analytical_database_b200 as adb
    if not os.path.exists(adb.lib_path()):
        adb.build_native()
    lib = C.CDLL(adb.lib_path())
    lib.adb_shared_select_plan.restype = C.c_int32
    lib.adb_shared_select_plan.argtypes = [I32, I32, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    return lib


class Plan:
    def __init__(self, lib, lows, highs):
        lows = np.ascontiguousarray(lows, dtype=np.int32)
        highs = np.ascontiguousarray(highs, dtype=np.int32)
        meta = (C.c_uint32 * 16)()
        assert lib.adb_shared_select_plan(lows.ctypes.data_as(I32), highs.ctypes.data_as(I32), lows.size,
                                          None, 0, meta) == 0
        buf = np.zeros(meta[0], dtype=np.uint8)
        assert lib.adb_shared_select_plan(lows.ctypes.data_as(I32), highs.ctypes.data_as(I32), lows.size,
                                          buf.ctypes.data_as(C.c_void_p), buf.size, meta) == 0
        (self.nbytes, self.m, self.lut_shift, self.bit_shift, self.span, lo, self.deepest, o_off, o_cov,
         o_lut, o_bits, o_q, o_cov4, n_lut, n_bits, q_stride) = [int(x) for x in meta]
        self.lo = int(np.int32(np.uint32(lo)))
        m = self.m
        self.bounds = buf[:4 * m].view(np.int32).astype(np.int64)
        self.cov_off = buf[o_off:o_off + 2 * (m + 2)].view(np.uint16).astype(np.int64)
        self.cov_q = buf[o_cov:o_cov + int(self.cov_off[-1]) if m else o_cov]
        self.lut = buf[o_lut:o_lut + 2 * (n_lut + 1)].view(np.uint16).astype(np.int64)
        self.bits = buf[o_bits:o_bits + n_bits // 8].view(np.uint32)
        self.q_first = buf[o_q:o_q + 2 * lows.size].view(np.uint16).astype(np.int64)
        self.q_last = buf[o_q + 2 * q_stride:o_q + 2 * q_stride + 2 * lows.size].view(np.uint16).astype(np.int64)
        self.cov4 = buf[o_cov4:o_cov4 + 4 * (m + 1)].view(np.uint32)
        self.n_lut, self.n_bits = n_lut, n_bits
        self.lows, self.highs = lows.astype(np.int64), highs.astype(np.int64)

    # the device's lookups, restated (shared_scan.cu: interval_of, the bitmap test)
    def interval_of(self, v):
        if self.m == 0 or v < self.lo:
            return 0
        if v >= self.lo + self.span:
            return self.m
        k = (v - self.lo) >> self.lut_shift
        i = int(self.lut[k])
        if i & 0x8000:
            return 0                                         # "no query reaches into this bucket"
        end = int(self.lut[k + 1]) & 0x7FFF
        while i < end and self.bounds[i] <= v:
            i += 1
        return i

    def maybe_hit(self, v):
        d = (v - self.lo) & 0xFFFFFFFF
        tb = min(d >> self.bit_shift, self.n_bits)
        if tb >= self.n_bits:
            return False
        return bool((int(self.bits[tb >> 5]) >> (tb & 31)) & 1)

    def cover(self, i):
        return [int(q) for q in self.cov_q[self.cov_off[i]:self.cov_off[i + 1]]]


def probe_values(rng, lows, highs):
    edge = np.concatenate([lows, highs, lows - 1, highs - 1, lows + 1, highs + 1]).astype(np.int64)
    lo, hi = int(edge.min()) - 1000, int(edge.max()) + 1000
    rnd = rng.integers(max(lo, -2**31), min(hi, 2**31 - 1), 4000)
    far = np.array([-2**31, 2**31 - 1, 0, -1, 1], dtype=np.int64)
    v = np.concatenate([edge, rnd, far])
    return v[(v >= -2**31) & (v < 2**31)]


def batches(rng):
    yield "disjoint", np.arange(0, 6000, 100), np.arange(0, 6000, 100) + 60
    yield "touching", np.arange(0, 6000, 100), np.arange(0, 6000, 100) + 100
    for depth in (2, 3, 4, 5, 9):
        lows = np.arange(0, 60) * 100
        yield f"staggered{depth}", lows, lows + 100 * depth
    yield "duplicates", np.array([5, 5, 5, 700, 700]), np.array([90, 90, 90, 800, 800])
    yield "nested150", np.arange(0, 150), 1000 - np.arange(0, 150)
    yield "empty+inverted", np.array([7, 50, 3]), np.array([7, 10, 4])
    yield "all empty", np.array([7, 50]), np.array([7, 10])
    yield "extremes", np.array([-2**31, -5, 2**31 - 100000, 0]), np.array([2**31 - 1, 5, 2**31 - 1, 2**30])
    lows = rng.integers(0, 100_000_000 - 100_000, 100)
    yield "config2", lows, lows + 100_000
    lows = rng.integers(-2**31, 2**31 - 2**20, 150)
    yield "wide random", lows, lows + rng.integers(0, 2**20, 150)
    lows = rng.integers(0, 200_000, 100)
    yield "m2 batch", lows, lows + rng.integers(0, 4000, 100)


def test_plan_tables_classify_like_the_reference_predicate(lib):
    rng = np.random.default_rng(7)
    for name, lows, highs in batches(rng):
        p = Plan(lib, lows, highs)
        live = p.lows < p.highs
        # bounds: the distinct bounds of the non-empty queries, ascending
        exp_bounds = np.unique(np.concatenate([p.lows[live], p.highs[live]])) if live.any() else np.array([], np.int64)
        assert np.array_equal(p.bounds, exp_bounds), name
        # cover lists per elementary interval, and the deepest cover
        depth = 1
        for k in range(1, p.m):
            a, b = p.bounds[k - 1], p.bounds[k]
            exp = [q for q in range(lows.size) if live[q] and p.lows[q] <= a and b <= p.highs[q]]
            assert p.cover(k) == exp, (name, k)
            depth = max(depth, len(exp))
        assert p.cover(0) == [] and (p.m == 0 or p.cover(p.m) == []), name
        assert p.deepest == depth, name
        # per-query interval ranges: exactly the intervals whose cover list names the query
        for q in range(lows.size):
            ids = [k for k in range(1, p.m) if q in p.cover(k)]
            if ids:
                assert (p.q_first[q], p.q_last[q]) == (ids[0], ids[-1]) and ids == list(range(ids[0], ids[-1] + 1))
            else:
                assert p.q_first[q] > p.q_last[q], (name, q)
        # every probe value: the prefilter never dismisses a hit, and the interval the tables
        # find is covered by exactly the queries whose predicate holds
        for v in probe_values(rng, p.lows, p.highs):
            v = int(v)
            want = [q for q in range(lows.size) if p.lows[q] <= v < p.highs[q]]
            if want:
                assert p.maybe_hit(v), (name, v)
            if p.maybe_hit(v):
                assert p.cover(p.interval_of(v)) == want, (name, v)
        # pair lists: overlapping queries never share a colour, every cover is there once
        if p.deepest <= 4:
            colour = {}
            for k in range(0, p.m + 1):
                word = int(p.cov4[k])
                qs = [(word >> (8 * c)) & 0xFF for c in range(4)]
                assert sorted(q for q in qs if q != 0xFF) == p.cover(k), (name, k)
                for c, q in enumerate(qs):
                    if q != 0xFF:
                        assert colour.setdefault(q, c) == c, (name, q)      # one colour per query


def test_plan_argument_checks(lib):
    meta = (C.c_uint32 * 16)()
    a = np.zeros(200, dtype=np.int32)
    p = a.ctypes.data_as(I32)
    assert lib.adb_shared_select_plan(p, p, 0, None, 0, meta) != 0          # q_count < 1
    assert lib.adb_shared_select_plan(p, p, 151, None, 0, meta) != 0        # server.c:366-371: chunks of 150
    assert lib.adb_shared_select_plan(None, p, 3, None, 0, meta) != 0
    assert lib.adb_shared_select_plan(p, p, 3, None, 0, meta) == 0
    small = np.zeros(16, dtype=np.uint8)
    assert lib.adb_shared_select_plan(p, p, 3, small.ctypes.data_as(C.c_void_p), small.size, meta) != 0
