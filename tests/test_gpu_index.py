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


def _view(eng, buf, offset_elems):
    from analytical_database_b200.engine import DevBuf
    v = DevBuf.__new__(DevBuf)
    v.eng, v.nbytes, v.ptr = eng, 0, buf.ptr + 4 * offset_elems
    return v


def test_full_size_index_properties(eng):
    """BASELINE config 3 at its full size: a 500 M-row column -- the index build and both range
    lookups, checked through size-independent properties (the oracle's quicksort needs
    minutes there): sortedness, permutation, value/position pairing, agreement of the index
    path with the scan path on counts and on the fetched values, B+-tree = sorted array."""
    n = 500_000_000
    key = eng.synth_uniform(n, 7, 0, 0, (1 << 31) - 1)
    pay = eng.synth_uniform(n, 8, 0, 0, 10000)
    vals, poss = eng.index_sort(key, n)
    # (1) values ascending: min(v[i+1] - v[i]) >= 0 (no wrap: values are in [0, 2^31))
    d = eng.aggregate(eng.ewise(_view(eng, vals, 1), vals, n - 1, True), n - 1)
    assert d.min >= 0
    # (2) pairing: values[i] == key[positions[i]] for every i
    f = eng.fetch(key, poss, n)
    z = eng.aggregate(eng.ewise(f, vals, n, True), n)
    assert z.min == 0 and z.max == 0
    f.free()
    # (3) positions are a permutation of 0 .. n-1: right sum, and -- sorted once more -- strictly
    #     ascending from 0 to n-1
    ps = eng.aggregate(poss, n)
    assert ps.sum == n * (n - 1) // 2 and ps.min == 0 and ps.max == n - 1
    sp, _ = eng.index_sort(poss, n)
    dd = eng.aggregate(eng.ewise(_view(eng, sp, 1), sp, n - 1, True), n - 1)
    assert dd.min == 1 and dd.max == 1
    sp.free(); _.free()
    # (4) stability (the order adb_index_sort documents): inside a run of equal values the
    #     positions ascend.  Equal neighbours have v[i+1] - v[i] == 0; there p[i+1] - p[i] > 0.
    #     Checked on a 4 M-entry window on the host.
    w = 4_000_000
    hv, hp = vals.to_host(w, offset_bytes=4 * 123_456_789), poss.to_host(w, offset_bytes=4 * 123_456_789)
    eq = hv[1:] == hv[:-1]
    assert np.all(hp[1:][eq] > hp[:-1][eq])
    # (5) range lookups against the scan path
    ix = eng.index_create(vals, poss, n, with_btree=True)
    vmin = int(vals.to_host(1)[0])
    for lo, width in [(max(vmin, 1000), (1 << 31) // 5000), (1 << 29, (1 << 31) // 100), (1 << 30, (1 << 31) // 10)]:
        hi = lo + width
        ps_, hs = eng.select_index_exact(ix, lo, hi, use_btree=False)
        pb_, hb = eng.select_index_exact(ix, lo, hi, use_btree=True)
        scan_pos, scan_h = eng.select_exact(key, n, lo, hi)
        assert hs == hb == scan_h
        same = eng.aggregate(eng.ewise(ps_, pb_, hs, True), hs)
        assert same.min == 0 and same.max == 0                       # B+-tree == sorted array
        fk = eng.aggregate(eng.fetch(key, ps_, hs), hs)
        assert lo <= fk.min and fk.max < hi                          # every emitted row qualifies
        # the same set of rows as the scan: equal position sums and equal payload aggregates
        a_ix, a_sc = eng.aggregate(ps_, hs), eng.aggregate(scan_pos, hs)
        assert (a_ix.sum, a_ix.min, a_ix.max) == (a_sc.sum, a_sc.min, a_sc.max)
        p_ix, p_sc = eng.aggregate(eng.fetch(pay, ps_, hs), hs), eng.aggregate(eng.fetch(pay, scan_pos, hs), hs)
        assert (p_ix.sum, p_ix.min, p_ix.max) == (p_sc.sum, p_sc.min, p_sc.max)
        for b in (ps_, pb_, scan_pos):
            b.free()
    eng.index_destroy(ix)
    for b in (key, pay, vals, poss):
        b.free()
