"""GPU parity for the index path (BASELINE config 3): sorted-array and B+-tree range select
against the oracle's restatement of select_column_sorted_index (query.c:165-198), on indexes
built by the oracle's exact Lomuto quicksort (index.c:25-46) so the tie order is the
reference's.  Covers the defined domain incl. the low == high quirk, the
oracle-undefined domain (low/high below the minimum, empty index, NULL bounds = scan
semantics), heavy duplicates (milestone3.py:46-56) and multi-level trees."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
I32MAX = 2**31 - 1


@pytest.fixture(scope="module")
def eng():
    from analytical_database_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def cases(rng):
    yield "perm5k", rng.permutation(5000).astype(np.int32)
    yield "dups", rng.integers(0, 50, 3000).astype(np.int32)
    yield "zipf", (rng.zipf(1.5, 4000) % 1000).astype(np.int32)
    yield "neg", rng.integers(-1000, 1000, 2000).astype(np.int32)
    yield "const", np.full(300, 7, np.int32)
    yield "one", np.array([5], np.int32)
    yield "n32", rng.integers(0, 10, 32).astype(np.int32)
    yield "n33", rng.integers(0, 10, 33).astype(np.int32)
    yield "n1025", rng.integers(-5, 40, 1025).astype(np.int32)
    yield "wide", rng.integers(-2**31, I32MAX, 40000, dtype=np.int64).astype(np.int32)


def probes(rng, values):
    vmin, vmax = int(values[0]), int(values[-1])
    out = {(vmin, vmin), (vmin, vmax), (vmin, min(vmax + 1, I32MAX)), (vmax, vmax),
           (vmax, min(vmax + 1, I32MAX)), (vmin, I32MAX), (vmin - 5 if vmin > -2**31 + 5 else vmin, vmax),
           (vmin - 9 if vmin > -2**31 + 9 else vmin, vmin - 2 if vmin > -2**31 + 9 else vmin)}
    for _ in range(60):
        lo = int(rng.integers(max(vmin - 3, -2**31), min(vmax + 4, I32MAX)))
        hi = int(rng.integers(max(vmin - 3, -2**31), min(vmax + 6, I32MAX)))
        out.add((lo, hi))
    return sorted(out)


def test_sorted_and_btree_select(eng, port, rng):
    for name, data in cases(rng):
        values, positions = port.index_sort(data)
        dv, dp = eng.upload(values), eng.upload(positions.astype(np.int32))
        ix = eng.index_create(dv, dp, data.size, with_btree=True)
        for lo, hi in probes(rng, values):
            exp, _undef = port.select_sorted_index(values, positions, lo, hi)
            for use_btree in (False, True):
                out, h = eng.select_index(ix, data.size, lo, hi, use_btree)
                assert h == exp.size, (name, lo, hi, use_btree)
                assert np.array_equal(out.to_host(h), exp), (name, lo, hi, use_btree)
                out.free()
        # NULL bounds: the reference dereferences NULL (query.c:208) -> scan semantics
        for lo, hi in [(None, None), (None, int(values[len(values) // 2])), (int(values[0]), None)]:
            exp = positions[(values >= (lo if lo is not None else -2**31)) &
                            (values < (hi if hi is not None else 2**31))].astype(np.int32)
            for use_btree in (False, True):
                out, h = eng.select_index(ix, data.size, lo, hi, use_btree)
                assert h == exp.size and np.array_equal(out.to_host(h), exp), (name, lo, hi)
        eng.index_destroy(ix)


def test_empty_index(eng):
    dv, dp = eng.upload(np.empty(0, np.int32)), eng.upload(np.empty(0, np.int32))
    ix = eng.index_create(dv, dp, 0, True)
    out, h = eng.select_index(ix, 0, 0, 5)
    assert h == 0
    eng.index_destroy(ix)


def test_index_select_matches_reference_objects(eng, ref, rng):
    """Same check against the UNMODIFIED reference function (defined domain only)."""
    data = rng.integers(0, 3000, 20000).astype(np.int32)
    values, positions = ref.index_sort(data)
    dv, dp = eng.upload(values), eng.upload(positions.astype(np.int32))
    ix = eng.index_create(dv, dp, data.size, True)
    vmin, vmax = int(values[0]), int(values[-1])
    for _ in range(80):
        lo = int(rng.integers(vmin, vmax + 2))
        hi = int(rng.integers(vmin, vmax + 3))
        exp, _ = ref.select_sorted_index(values, positions, lo, hi)
        for use_btree in (False, True):
            out, h = eng.select_index(ix, data.size, lo, hi, use_btree)
            assert h == exp.size and np.array_equal(out.to_host(h), exp), (lo, hi, use_btree)
            out.free()
    eng.index_destroy(ix)


def test_large_tree_levels(eng, port, rng):
    """3.4 M unique keys -> 4 tree levels; result checked against numpy searchsorted and the
    scan (set-equal to the index result when keys are unique)."""
    n = 3_400_001
    data = rng.permutation(n).astype(np.int32)
    positions = np.argsort(data, kind="stable").astype(np.int32)
    values = data[positions]
    col = eng.upload(data)
    dv, dp = eng.upload(values), eng.upload(positions)
    ix = eng.index_create(dv, dp, n, True)
    for lo, hi in [(0, 1), (5, 70000), (1_000_000, 1_340_000), (n - 10, n + 5), (n, n + 1), (123, 123)]:
        exp, _ = port.select_sorted_index(values, positions.astype(np.uint64), lo, hi)
        for use_btree in (False, True):
            out, h = eng.select_index(ix, n, lo, hi, use_btree)
            got = out.to_host(h)
            assert h == exp.size and np.array_equal(got, exp), (lo, hi, use_btree)
            out.free()
        if lo < hi:
            spos, _c, sh = eng.select_scan(col, n, lo, hi)
            assert np.array_equal(np.sort(got), spos.to_host(sh))
            spos.free()
    eng.index_destroy(ix)
