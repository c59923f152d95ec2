#!/usr/bin/env python
"""Device-resident chain (adb_chain_select_fetch_agg) over 500 M-row shards for several slice
settings (adb_chain_config): ms per shard, CUDA events.  Development probe."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import analytical_database_b200 as adb  # noqa: E402

eng = adb.Engine(0)
n = 500_000_000
span = 1 << 30
shards = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cols = [(eng.synth_uniform(n, 42, s * n, 0, span), eng.synth_uniform(n, 43, s * n, 2**31 - 10000, 10000)) for s in range(shards)]
out = {}
for sel in (0.01, 0.1):
    lo, hi = 1000, 1000 + int(span * sel)
    cap = int(n * sel * 1.2) + 4096
    res = [(eng.alloc_i32(cap), eng.alloc_i32(cap), eng.alloc(8), eng.alloc(64)) for _ in range(shards)]
    blo, bhi = C.c_int32(lo), C.c_int32(hi)
    for slices, div in [(1, 2), (2, 2), (4, 2), (4, 1), (4, 4), (8, 2), (8, 4), (16, 4), (6, 3)]:
        eng._ck(eng.lib.adb_chain_config(slices, div))

        def step():
            for (c1, c2), (p, v, cnt, agg) in zip(cols, res):
                eng._ck(eng.lib.adb_chain_select_fetch_agg(c1.i32(), c2.i32(), n, C.byref(blo), C.byref(bhi),
                                                           p.i32(), v.i32(), cnt.i64(), eng.agg_ptr(agg)))
        for _ in range(3):
            step()
        eng.sync()
        eng.timer_start()
        for _ in range(10):
            step()
        ms = eng.timer_stop() / 10 / shards
        a = eng.read_agg(res[0][3])
        out[f"sel{sel}_s{slices}_d{div}"] = {"ms_per_shard": round(ms, 4), "sum": a.sum, "count": a.count,
                                            "frac_of_peak": round((4.0 * n + 20.0 * a.count) / (ms * 1e-3) / 1e9 / 6550.7, 4)}
    for r in res:
        for b in r:
            b.free()
print(json.dumps(out, indent=1))
