#!/usr/bin/env python
"""Kernel times (CUDA events) of the two halves of the lazy chain on one GPU: the predicate pass
(adb_select_count_base, device count only) and the aggregate-only resolution
(adb_select_emit_fetch_agg with NULL outputs), plus the one-kernel form.  Development probe."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import analytical_database_b200 as adb  # noqa: E402

eng = adb.Engine(0)
span = 1 << 30
out = {}
for n in (250_000_000, 500_000_000):
    c1 = eng.synth_uniform(n, 42, 0, 0, span)
    c2 = eng.synth_uniform(n, 43, 0, 2**31 - 10000, 10000)
    cnt, agg = eng.alloc(8), eng.alloc(64)
    for sel in (0.001, 0.01, 0.1):
        blo, bhi = C.c_int32(1000), C.c_int32(1000 + int(span * sel))

        def count():
            eng._ck(eng.lib.adb_select_count_base(c1.i32(), n, C.byref(blo), C.byref(bhi), 0, cnt.i64(), None))

        def fold():
            eng._ck(eng.lib.adb_select_emit_fetch_agg(c2.i32(), None, None, eng.agg_ptr(agg), None))

        def one():
            eng._ck(eng.lib.adb_chain_select_agg(c1.i32(), c2.i32(), n, C.byref(blo), C.byref(bhi), cnt.i64(),
                                                 eng.agg_ptr(agg), None))
        res = {}
        for name, fn, pre in (("count", count, None), ("fold", fold, count), ("one_kernel", one, None)):
            if pre:
                pre()
            for _ in range(3):
                fn()
            eng.sync()
            eng.timer_start()
            for _ in range(20):
                fn()
            res[name + "_us"] = round(eng.timer_stop() / 20 * 1e3, 1)
        a = eng.read_agg(agg)
        res["hits"] = a.count
        out[f"n{n}_sel{sel}"] = res
    for b in (c1, c2, cnt, agg):
        b.free()
print(json.dumps(out, indent=1))
