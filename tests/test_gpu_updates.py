"""Updates and deletes: relational_update / relational_delete of milestone 5.

The reference does not implement them (its parser has no branch, src/parse.c:876-960), so there
is no reference function to diff against; the oracle is the model the reference's own generator
uses to compute its expected outputs (project_tests/data_generation_scripts/milestone5.py:123-262:
`dataTable.loc[mask, 'col1'] = v` for an update, `dataTable = dataTable[dataTable.col != v]` for a
delete, rows keep their order), restated with numpy.  Every query after a change is also checked
against the reference's own select / fetch on the model's arrays."""
import ctypes as C

import numpy as np
import pytest

from query_api import Api, Column

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


def colptrs(cols):
    return (C.POINTER(Column) * len(cols))(*[C.pointer(c) for c, _ in cols])


def model_of(cols):
    return [a[:c.row_count].copy() for c, a in cols]


def check_queries(api, cpu, cols, model, rng):
    n = model[0].size
    for _ in range(3):
        j, k = rng.integers(0, len(cols), 2)
        lo, hi = sorted(int(x) for x in rng.integers(-300, 1200, 2))
        s = api.select_column(cols[j][0], lo, hi)
        f = api.fetch_column(cols[k][0], s)
        epos = cpu.select_scan(model[j], lo, hi)
        assert np.array_equal(api.tuples(s), epos)
        assert np.array_equal(api.tuples(f), cpu.fetch(model[k], epos))
        api.drop(s), api.drop(f)
    for j, (c, a) in enumerate(cols):                          # the catalog's host arrays follow
        assert c.row_count == n and np.array_equal(a[:n], model[j])


def test_updates_and_deletes_follow_the_generator_model(api, cpu, rng):
    """milestone5.py tests 40-42 in spirit: updates by predicate, deletes by predicate, queries in
    between, on a 4-column table."""
    n = 50_000
    arrays = [rng.integers(0, 1000, n).astype(np.int32), rng.integers(0, 1000, n).astype(np.int32),
              rng.integers(0, 10000, n).astype(np.int32), rng.integers(0, 10000, n).astype(np.int32)]
    cols = api.table(arrays)
    model = model_of(cols)
    ptrs = colptrs(cols)
    # UPDATE tbl SET col1 = v WHERE colK in [lo, hi)   (milestone5.py:123-160)
    for k, lo, hi, v in [(0, 10, 11, -10), (1, 22, 23, -20), (0, 300, 340, -30), (2, 4440, 4500, -40), (0, -10, -9, -50)]:
        u = api.select_column(cols[k][0], lo, hi)
        assert api.lib.adb_host_relational_update(ptrs, len(cols), 0, u, v) == 0, api.lib.adb_host_last_error()
        mask = (model[k] >= lo) & (model[k] < hi)
        model[0][mask] = v
        api.drop(u)
        check_queries(api, cpu, cols, model, rng)
    assert cols[0][0].min == -50                               # min / max follow the updates
    # DELETE FROM tbl WHERE colK in [lo, hi)            (milestone5.py:176-214)
    for k, lo, hi in [(0, -50, -49), (1, 22, 23), (0, -30, -29), (2, 4440, 4500), (3, 0, 2500), (0, 5000, 6000)]:
        d = api.select_column(cols[k][0], lo, hi)
        assert api.lib.adb_host_relational_delete(ptrs, len(cols), d) == 0, api.lib.adb_host_last_error()
        keep = ~((model[k] >= lo) & (model[k] < hi))
        model = [m[keep] for m in model]
        api.drop(d)
        check_queries(api, cpu, cols, model, rng)
    assert model[0].size < n


def test_handles_created_before_a_change_keep_their_values(api, cpu, rng):
    n = 20_011
    arrays = [rng.integers(0, 100, n).astype(np.int32), rng.integers(-1000, 1000, n).astype(np.int32)]
    cols = api.table(arrays)
    model = model_of(cols)
    ptrs = colptrs(cols)
    s = api.select_column(cols[0][0], 10, 40)                  # unwritten (lazy) ...
    f = api.fetch_column(cols[1][0], s)                        # ... and its unwritten fetch
    epos = cpu.select_scan(model[0], 10, 40)
    evals = cpu.fetch(model[1], epos)
    u = api.select_column(cols[0][0], 20, 30)
    assert api.lib.adb_host_relational_update(ptrs, 2, 1, u, 7777) == 0
    assert np.array_equal(api.tuples(f), evals)                # pre-update values
    assert np.array_equal(api.tuples(s), epos)
    assert api.lib.adb_host_relational_delete(ptrs, 2, u) == 0
    assert np.array_equal(api.tuples(f), evals) and np.array_equal(api.tuples(s), epos)
    for r in (s, f, u):
        api.drop(r)
    keep = ~((model[0] >= 20) & (model[0] < 30))
    model[1][(model[0] >= 20) & (model[0] < 30)] = 7777
    model = [m[keep] for m in model]
    check_queries(api, cpu, cols, model, rng)


def test_positions_in_any_order_and_an_unclustered_index(api, cpu, rng):
    """Positions from foreign code (any order, duplicates) and a column with an unclustered
    index, which is rebuilt on the engine after every change."""
    n = 30_000
    key = rng.permutation(n).astype(np.int32)
    pay = rng.integers(0, 500, n).astype(np.int32)
    cols = api.table([key, pay], {0: (True, False)})
    ptrs = colptrs(cols)
    api.build_index(cols, 0)
    model = model_of(cols)
    pos = rng.integers(0, n, 900).astype(np.int32)             # duplicates, no order
    hp = api.host_result(pos)
    assert api.lib.adb_host_relational_update(ptrs, 2, 1, C.pointer(hp), -5) == 0, api.lib.adb_host_last_error()
    model[1][pos] = -5
    dele = rng.choice(n, 4000, replace=False).astype(np.int32)
    hd = api.host_result(dele)
    assert api.lib.adb_host_relational_delete(ptrs, 2, C.pointer(hd)) == 0, api.lib.adb_host_last_error()
    keep = np.ones(n, bool)
    keep[dele] = False
    model = [m[keep] for m in model]
    m = model[0].size
    assert cols[0][0].row_count == m
    # index select (value order) on the rebuilt index, then fetch
    ev, ep = cpu.index_sort(model[0])
    lo, hi = int(ev[m // 4]), int(ev[m // 2])
    s = api.select_column(cols[0][0], lo, hi)
    exp, undefined = cpu.select_sorted_index(ev, ep, lo, hi)
    assert not undefined and np.array_equal(api.tuples(s), exp)
    f = api.fetch_column(cols[1][0], s)
    assert np.array_equal(api.tuples(f), model[1][exp])
    api.drop(s), api.drop(f)
    # updating the indexed key itself rebuilds its index
    u = api.select_column(cols[1][0], -5, -4)
    assert api.lib.adb_host_relational_update(ptrs, 2, 0, u, -77) == 0, api.lib.adb_host_last_error()
    model[0][model[1] == -5] = -77
    api.drop(u)
    s = api.select_column(cols[0][0], -77, -76)
    assert np.array_equal(np.sort(api.tuples(s)), np.flatnonzero(model[0] == -77).astype(np.int32))
    api.drop(s)


def test_refusals(api, rng):
    n = 1000
    cols = api.table([rng.permutation(n).astype(np.int32), np.arange(n, dtype=np.int32)], {0: (True, True)})
    ptrs = colptrs(cols)
    api.build_index(cols, 0)
    hp = api.host_result(np.arange(5, dtype=np.int32))
    assert api.lib.adb_host_relational_delete(ptrs, 2, C.pointer(hp)) == -1
    assert b"clustered" in api.lib.adb_host_last_error()
    assert api.lib.adb_host_relational_update(ptrs, 2, 1, C.pointer(hp), 3) == -1
