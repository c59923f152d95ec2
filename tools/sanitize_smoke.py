#!/usr/bin/env python
"""Every kernel of the engine once, at small sizes, with results checked against the oracle:
meant to run under compute-sanitizer (memcheck / racecheck / synccheck), which is far too
slow for the full GPU suite.   compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import analytical_database_b200 as adb  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    import ctypes as C
    eng, port = adb.Engine(0), oracle.port()
    rng = np.random.default_rng(3)
    n = 70_003
    a = rng.integers(-5000, 5000, n).astype(np.int32)
    b = rng.integers(2**31 - 10000, 2**31 - 1, n, dtype=np.int64).astype(np.int32)
    da, db = eng.upload(a), eng.upload(b)
    for lo, hi in [(-100, 900), (None, None), (7, 8), (4000, None)]:
        pos, dc, h = eng.select_scan(da, n, lo, hi)
        e = port.select_scan(a, lo, hi)
        assert h == e.size and np.array_equal(pos.to_host(h), e)
        p2, h2 = eng.select_exact(da, n, lo, hi)
        assert h2 == h and np.array_equal(p2.to_host(h), e)
        val = eng.fetch(db, pos, h)
        assert np.array_equal(val.to_host(h), port.fetch(b, e))
        g = eng.aggregate(val, h)
        assert g.sum == port.sum(port.fetch(b, e)) and g.count == h
        sp, sdc, sh = eng.select_pairs(val, pos, h, 2**31 - 6000, None)
        assert np.array_equal(sp.to_host(sh), port.select_result(port.fetch(b, e), e, 2**31 - 6000, None))
        s = eng.ewise(val, val, h, True)
        assert not s.to_host(h).any()
    # fused chain
    pos, val, dcnt, dagg = eng.alloc_i32(n), eng.alloc_i32(n), eng.alloc(8), eng.alloc(64)
    blo, bhi = C.c_int32(-2000), C.c_int32(3000)
    eng._ck(eng.lib.adb_chain_select_fetch_agg(da.i32(), db.i32(), n, C.byref(blo), C.byref(bhi), pos.i32(),
                                               val.i32(), dcnt.i64(), eng.agg_ptr(dagg)))
    e = port.select_scan(a, -2000, 3000)
    g = eng.read_agg(dagg)
    assert g.count == e.size and g.sum == port.sum(port.fetch(b, e))
    assert np.array_equal(val.to_host(e.size), port.fetch(b, e))
    # deferred emit (count, host sizes both lists, positions + gather + aggregates in one kernel)
    dp, dv, dh, dg = eng.select_fetch_agg_deferred(da, db, n, -2000, 3000)
    assert dh == e.size and (dg.sum, dg.count, dg.min, dg.max) == (g.sum, g.count, g.min, g.max)
    assert np.array_equal(dp.to_host(dh), e) and np.array_equal(dv.to_host(dh), port.fetch(b, e))
    # shared scan: sparse, dense and nested batches
    for lows, highs in [(rng.integers(-5000, 4000, 40), None), (np.arange(0, 150), 1000 - np.arange(0, 150))]:
        lows = lows.astype(np.int32)
        highs = (lows + 50).astype(np.int32) if highs is None else highs.astype(np.int32)
        got = eng.shared_select(da, n, lows, highs)
        exp = port.shared_select(a, lows, highs)
        for (buf, c), x in zip(got, exp):
            assert c == x.size and np.array_equal(buf.to_host(c), x)
    # index: sort, create (with B+-tree), select both ways
    vals, poss = eng.index_sort(da, n)
    order = np.argsort(a, kind="stable")
    assert np.array_equal(vals.to_host(n), a[order]) and np.array_equal(poss.to_host(n), order.astype(np.int32))
    ix = eng.index_create(vals, poss, n, True)
    for tree in (False, True):
        p, c = eng.select_index_exact(ix, -100, 900, use_btree=tree)
        assert np.array_equal(np.sort(p.to_host(c)), port.select_scan(a, -100, 900))
    eng.index_destroy(ix)
    # joins + routing
    k1 = rng.integers(1, 3000, 20_000).astype(np.int32)
    k2 = rng.integers(1, 3000, 9_000).astype(np.int32)
    q1, q2 = rng.permutation(20_000).astype(np.int32), rng.permutation(9_000).astype(np.int32)
    d = [eng.upload(x) for x in (k1, q1, k2, q2)]
    for nested in (False, True):
        o1, o2, m = eng.join(d[0], d[1], 20_000, d[2], d[3], 9_000, nested_loop=nested)
        e1, e2 = (port.nested_loop_join if nested else port.hash_join)(k1, q1, k2, q2)
        assert m == e1.size and np.array_equal(o1.to_host(m), e1) and np.array_equal(o2.to_host(m), e2)
    ov, op = eng.alloc_i32(20_000), eng.alloc_i32(20_000)
    counts = (C.c_int64 * 8)()
    eng._ck(eng.lib.adb_route_pairs(d[0].i32(), d[1].i32(), 20_000, 8, ov.i32(), op.i32(), counts))
    assert sum(counts) == 20_000
    # a skewed join: one heavy key overflows the shared-memory table (global-slot path)
    k3 = np.where(rng.random(12_000) < 0.6, 7, rng.integers(1, 500, 12_000)).astype(np.int32)
    q3 = np.arange(12_000, dtype=np.int32)
    d3 = [eng.upload(x) for x in (k3, q3)]
    o1, o2, m = eng.join(d3[0], d3[1], 12_000, d[2], d[3], 9_000)
    e1, e2 = port.hash_join(k3, q3, k2, q2)
    assert m == e1.size and np.array_equal(o1.to_host(m), e1) and np.array_equal(o2.to_host(m), e2)
    # bulk load + result text
    text = ("h\n" + "\n".join(",".join(str(int(x)) for x in r) for r in rng.integers(-10**6, 10**6, (3001, 3))) +
            "\n 7,abc\n\n9223372036854775808,-5,+3,99").encode()
    cols, rows = eng.csv_load(text, 3)
    exp = port.csv_parse(text, 3)
    assert rows == exp.shape[1]
    for c, x in zip(cols, exp):
        assert np.array_equal(c.to_host(rows), x)
    assert eng.format_i32(cols[0], rows) == port.print_i32(exp[0])
    # the exchange-carrying chain kernel and the stand-alone exchange, world size 1
    hbuf = C.create_string_buffer(64)
    eng._ck(eng.lib.adb_peer_create(1, 0, hbuf))
    eng._ck(eng.lib.adb_peer_connect(hbuf.raw))
    dout = eng.alloc(64)
    for _ in range(3):
        eng._ck(eng.lib.adb_chain_select_fetch_agg_exchange(da.i32(), db.i32(), n, C.byref(blo), C.byref(bhi),
                                                            pos.i32(), val.i32(), dcnt.i64(), eng.agg_ptr(dagg), 1,
                                                            eng.agg_ptr(dout)))
        g2 = eng.read_agg(dout)
        assert (g2.sum, g2.count, g2.min, g2.max) == (g.sum, g.count, g.min, g.max)
    eng._ck(eng.lib.adb_agg_combine_allreduce(eng.agg_ptr(dagg), 1, eng.agg_ptr(dout), None))
    g2 = eng.read_agg(dout)
    assert (g2.sum, g2.count) == (g.sum, g.count)
    eng._ck(eng.lib.adb_peer_destroy())
    eng.sync()
    print("sanitize_smoke ok:", eng.launch_count(), "launches")
    eng.close()


if __name__ == "__main__":
    main()
