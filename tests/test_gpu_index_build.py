"""GPU parity of the index build: adb_host_index_build of the C host
shim -- radix sort on the engine, clustered sibling permutation, host catalog arrays filled --
against the reference's own index.c (oracle/_ref: quicksort / init_column_index /
reorder_column, src/index.c:25-146).  Unique keys: bit-exact.  Duplicate keys: the reference's
tie order is that of its unstable quicksort; the engine's is ascending row order,
so values must be identical and positions equal as a set per key."""
import ctypes as C

import numpy as np
import pytest

from query_api import Api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 3], ids=["1gpu", "3gpus"])
def api(request):
    a = Api()
    assert a.lib.adb_host_init_multi(request.param) == 0, a.lib.adb_host_last_error()
    yield a
    a.lib.adb_host_shutdown()


@pytest.fixture(scope="module")
def cpu():
    from oracle import oracle
    return oracle.reference("O2") or oracle.port()


@pytest.mark.parametrize("n", [1, 33, 5000, 300_007])
@pytest.mark.parametrize("kind", ["sorted", "btree"])
def test_unclustered_unique_keys_equal_the_reference_build(api, cpu, rng, n, kind):
    key = (rng.permutation(n) * 7 - n).astype(np.int32)                  # unique, negative to positive
    pay = rng.integers(-10**6, 10**6, n).astype(np.int32)
    cols = api.table([pay, key], {1: (kind == "sorted", False)})
    values, positions = api.build_index(cols, 1)
    ev, ep = cpu.index_sort(key)                                        # src/index.c:25-46,140-146
    assert np.array_equal(values, ev) and np.array_equal(positions, ep)
    assert np.array_equal(cols[0][1], pay) and np.array_equal(cols[1][1], key)   # nothing moves
    # the device-side index installed by the build answers selects (no re-upload) like the reference
    for lo, hi in [(int(ev[0]), int(ev[0]) + 1), (int(ev[n // 3]), int(ev[(2 * n) // 3]) + 1), (5, 5),
                   (int(ev[-1]), int(ev[-1]) + 50)]:
        if lo > hi:
            continue
        s = api.select_column(cols[1][0], lo, hi)
        exp, undefined = cpu.select_sorted_index(ev, ep, lo, hi)
        assert not undefined and np.array_equal(api.tuples(s), exp), (lo, hi)
        f = api.fetch_column(cols[0][0], s)
        assert np.array_equal(api.tuples(f), pay[exp])
        api.drop(s), api.drop(f)


@pytest.mark.parametrize("n", [2, 4099, 200_003])
def test_clustered_build_permutes_the_siblings_only(api, cpu, rng, n):
    """index.c:119-135: the indexed column keeps its load-order data and identity positions,
    every other column moves into index order."""
    key = rng.permutation(n).astype(np.int32)
    a, b = rng.integers(0, 1000, n).astype(np.int32), rng.integers(-5, 5, n).astype(np.int32)
    cols = api.table([a, key, b], {1: (True, True)})
    values, positions = api.build_index(cols, 1)
    ev, ep = cpu.index_sort(key)
    assert np.array_equal(values, ev)
    assert np.array_equal(positions, np.arange(n, dtype=np.uint64))      # identity
    assert np.array_equal(cols[1][1], key)                               # untouched
    assert np.array_equal(cols[0][1], cpu.reorder(a, ep)) and np.array_equal(cols[2][1], cpu.reorder(b, ep))
    # select on the clustered column -> fetch of a sibling is right; -> fetch of itself returns
    # load-order data, as in the reference
    lo, hi = n // 4, n // 2 + 1
    s = api.select_column(cols[1][0], lo, hi)
    exp, undefined = cpu.select_sorted_index(ev, np.arange(n, dtype=np.uint64), lo, hi)
    assert not undefined and np.array_equal(api.tuples(s), exp)
    fa, fk = api.fetch_column(cols[0][0], s), api.fetch_column(cols[1][0], s)
    assert np.array_equal(api.tuples(fa), cpu.reorder(a, ep)[exp])
    assert np.array_equal(api.tuples(fk), key[exp])
    for r in (s, fa, fk):
        api.drop(r)


def test_declaration_order_quirk(api, cpu, rng):
    """build_index walks the columns in declaration order (index.c:158-175): an unclustered index
    on an earlier column is built on pre-permutation rows and keeps those positions when a later
    column's clustered index permutes the table (tbl4 of milestone 3; test 25 fails for it)."""
    n = 20_011
    c1 = rng.integers(0, 10**6, n).astype(np.int32)
    c2 = rng.permutation(n).astype(np.int32)
    c3 = (rng.permutation(n) - 500).astype(np.int32)
    cols = api.table([c1, c2, c3], {1: (False, False), 2: (True, True)})
    v2, p2 = api.build_index(cols, 1)
    v3, p3 = api.build_index(cols, 2)
    e2v, e2p = cpu.index_sort(c2)
    e3v, e3p = cpu.index_sort(c3)
    assert np.array_equal(v2, e2v) and np.array_equal(p2, e2p)
    assert np.array_equal(v3, e3v) and np.array_equal(p3, np.arange(n, dtype=np.uint64))
    assert np.array_equal(cols[0][1], cpu.reorder(c1, e3p)) and np.array_equal(cols[1][1], cpu.reorder(c2, e3p))
    # select on col2 (stale positions) -> fetch col1 (permuted): what the reference would return
    s = api.select_column(cols[1][0], 100, 300)
    exp, undefined = cpu.select_sorted_index(e2v, e2p, 100, 300)
    assert not undefined and np.array_equal(api.tuples(s), exp)
    f = api.fetch_column(cols[0][0], s)
    assert np.array_equal(api.tuples(f), cpu.reorder(c1, e3p)[exp])
    api.drop(s), api.drop(f)


@pytest.mark.parametrize("n", [1000, 150_001])
def test_duplicate_keys_stable_tie_order(api, cpu, rng, n):
    key = rng.integers(0, 300, n).astype(np.int32)                       # ~n/300 rows per key
    sib = np.arange(n, dtype=np.int32)
    cols = api.table([sib, key], {1: (True, False)})
    values, positions = api.build_index(cols, 1)
    ev, ep = cpu.index_sort(key)
    assert np.array_equal(values, ev)
    order = np.argsort(key, kind="stable")
    assert np.array_equal(positions, order.astype(np.uint64))            # ties: ascending row order
    # the same tuples as the reference's build, key by key
    assert np.array_equal(np.sort(positions.reshape(-1)), np.sort(ep))
    assert np.array_equal(key[positions.astype(np.int64)], key[ep.astype(np.int64)])
    s = api.select_column(cols[1][0], 17, 42)
    exp, undefined = cpu.select_sorted_index(ev, ep, 17, 42)
    got = api.tuples(s)
    assert not undefined and np.array_equal(np.sort(got), np.sort(exp))
    assert np.array_equal(key[got], key[exp])                            # same value order
    api.drop(s)


def test_histogram_counts(api, rng):
    """build_histogram, index.c:63-84: 100 bins of width (max - min) / 99 from column->min."""
    n = 123_457
    data = rng.integers(-40_000, 60_000, n).astype(np.int32)
    col = api.column(data)
    bin_size = (int(data.max()) - int(data.min())) // 99
    counts = (C.c_ulong * 100)()
    assert api.lib.adb_host_column_histogram(C.byref(col), bin_size, counts) == 0
    bins = (data.astype(np.int64) - int(data.min())) // bin_size
    exp = np.bincount(bins[bins < 100], minlength=100)
    assert np.array_equal(np.array(counts[:], dtype=np.int64), exp)


def test_config3_full_size_through_select_column(rng):
    """BASELINE config 3 at full size through the operator API: 500 M rows, indexed key = a
    permutation of 0..n-1 ((row * mul + add) mod n: unique keys, so every correct sort is the
    reference's), payload uniform in [0, 10000).  Index build on the engine,
    btree range select + fetch at 0.02 % and 1 %; oracle = the closed form of the permutation
    (the row holding key v is inv(mul) * (v - add) mod n), which
    test_unclustered_unique_keys_equal_the_reference_build pins to the reference's own index build
    at sizes it can sort.  The same run on 2 engine contexts (sliced index, peer gathers)."""
    import ctypes as C
    import analytical_database_b200 as adb
    from analytical_database_b200 import synth
    import query_api as q
    n, mul, add = 500_000_000, 387_420_489, 123_456_789
    inv = pow(mul, -1, n)
    for G in (1, 2):
        api = Api()
        assert api.lib.adb_host_init_multi(G) == 0, api.lib.adb_host_last_error()
        eng = adb.Engine(0)
        S = ((n + G - 1) // G + 31) // 32 * 32
        cols = []
        for name, fill in (("key", lambda first, cnt: _affine(eng, cnt, first, mul, add, n)),
                           ("pay", lambda first, cnt: eng.synth_uniform(cnt, 8, first, 0, 10000))):
            ptrs, bufs = (C.c_void_p * G)(), []
            for g in range(G):
                eng._ck(eng.lib.adb_ctx_select(g))
                b = fill(g * S, max(1, min(S, n - g * S)))
                eng.sync()
                bufs.append(b)
                ptrs[g] = b.ptr
            eng._ck(eng.lib.adb_ctx_select(0))
            c = q.Column()
            c.name, c.row_count = name.encode(), n
            if name == "key":
                c.has_index, c.sorted, c.clustered = True, False, False
            assert api.lib.adb_host_column_adopt_shards(C.byref(c), ptrs, S) == 0
            cols.append((c, bufs))
        (key, _), (pay, _) = cols
        arr = (C.POINTER(q.Column) * 2)(C.pointer(key), C.pointer(pay))
        assert api.lib.adb_host_index_build(arr, 2, 0) == 0, api.lib.adb_host_last_error()
        hv = np.ctypeslib.as_array(key.index.contents.values, shape=(n,))
        hp = np.ctypeslib.as_array(key.index.contents.positions, shape=(n,))
        for a0 in (0, n // 3, n - (1 << 18)):                              # the catalog's host arrays
            vs = np.arange(a0, a0 + (1 << 18), dtype=np.int64)
            assert np.array_equal(hv[a0:a0 + (1 << 18)], vs.astype(np.int32))
            assert np.array_equal(hp[a0:a0 + (1 << 18)], ((vs - add) % n * inv % n).astype(np.uint64))
        for sel in (0.0002, 0.01):
            lo = n // 7
            hi = lo + int(n * sel)
            s_ = api.select_column(key, lo, hi)
            f_ = api.fetch_column(pay, s_)
            assert s_.contents.num_tuples == hi - lo
            gpos, gval = api.tuples(s_), api.tuples(f_)
            vs = np.arange(lo, hi, dtype=np.int64)
            epos = ((vs - add) % n * inv % n).astype(np.int32)
            z = synth.mix64(8, epos.astype(np.uint64))
            evals = (((z >> np.uint64(32)) * np.uint64(10000)) >> np.uint64(32)).astype(np.int32)
            assert np.array_equal(gpos, epos) and np.array_equal(gval, evals), (G, sel)
            api.drop(s_), api.drop(f_)
        ix = key.index.contents
        for ptr in (ix.values, ix.positions, key.index):
            q._libc.free(C.cast(ptr, C.c_void_p))
        api.lib.adb_host_shutdown()


def _affine(eng, cnt, first, mul, add, n):
    b = eng.alloc_i32(cnt)
    eng._ck(eng.lib.adb_synth_affine(b.i32(), cnt, first, mul, add, n))
    return b
