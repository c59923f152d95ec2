#!/usr/bin/env python
"""BASELINE config 1 through the unchanged client/server: a 1 M-row, 4-column int table shaped
like project_tests milestone 1 (col1, col2 in [-N/2, N/2), col3 in [0, 100), col4 in
[2^31-10000, 2^31); milestone1.py:112-121), loaded from CSV, then the milestone-1 query
forms (select range + fetch + sum / avg / min / max / add / sub, chained select, a batch)
replayed through BOTH pairs built by oracle/Makefile:

    server_ref  + client_ref     the unmodified reference, operators on the host CPU
    server_b200 + client_b200    the same plumbing linked with host/query_shim.c + libadb_b200.so

Checks the two clients print identical bytes and reports wall-clock per DSL script (client
start to client exit: socket + parse + operators + print).  Development / evidence tool:
    python tools/dropin_c1.py [rows=1000000] > profiles/r01_dropin_c1.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dsl_harness as H  # noqa: E402


def scripts(n, csv):
    q = n // 4
    return {
        "load": f'create(db,"db1")\ncreate(tbl,"tbl2",db1,4)\ncreate(col,"col1",db1.tbl2)\n'
                f'create(col,"col2",db1.tbl2)\ncreate(col,"col3",db1.tbl2)\ncreate(col,"col4",db1.tbl2)\n'
                f'load("{csv}")\n',
        "select_fetch_sum": f"s1=select(db1.tbl2.col1,{-q},{q})\nf1=fetch(db1.tbl2.col3,s1)\na1=sum(f1)\nprint(a1)\n"
                            "a2=sum(db1.tbl2.col1)\nprint(a2)\n",
        "select_fetch_avg": f"s1=select(db1.tbl2.col1,{-q // 10},{q // 10})\nf1=fetch(db1.tbl2.col3,s1)\n"
                            "a1=avg(f1)\nprint(a1)\n",
        "min_max": f"s1=select(db1.tbl2.col1,{-q},null)\nf1=fetch(db1.tbl2.col2,s1)\nm1=min(f1)\nm2=max(f1)\n"
                   "print(m1,m2)\n",
        "add_sub_print_rows": f"s11=select(db1.tbl2.col1,{-q},{-q + 40})\nf11=fetch(db1.tbl2.col2,s11)\n"
                              "f12=fetch(db1.tbl2.col3,s11)\na11=add(f11,f12)\nprint(a11)\ns21=sub(f12,f11)\nprint(s21)\n",
        "chained_select": f"s1=select(db1.tbl2.col1,{-q},{q})\nsf1=fetch(db1.tbl2.col2,s1)\n"
                          f"s2=select(s1,sf1,{-q // 2},{q // 2})\nf1=fetch(db1.tbl2.col1,s2)\nf2=fetch(db1.tbl2.col2,s2)\n"
                          "f3=fetch(db1.tbl2.col3,s2)\nadd12=add(f1,f2)\nout1=avg(add12)\nout2=min(f2)\nout3=max(f3)\n"
                          "sub32=sub(f3,f2)\nout4=avg(sub32)\nout5=sum(sub32)\nprint(out1,out2,out3,out4,out5)\n",
        "batch_of_20": "batch_queries()\n" + "".join(
            f"b{i}=select(db1.tbl2.col1,{-q + i * 1000},{-q + i * 1000 + 500})\n" for i in range(20)) +
            "batch_execute()\n" + "".join(f"g{i}=fetch(db1.tbl2.col3,b{i})\nh{i}=sum(g{i})\nprint(h{i})\n"
                                          for i in range(20)),
    }


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rng = np.random.default_rng(42)
    work = tempfile.mkdtemp(prefix="adb_c1_")
    csv = os.path.join(work, "data2.csv")
    tab = np.stack([rng.integers(-n // 2, n // 2, n), rng.integers(-n // 2, n // 2, n),
                    rng.integers(0, 100, n), rng.integers(2**31 - 10000, 2**31, n)], 1)
    np.savetxt(csv, tab, fmt="%d", delimiter=",",
               header="db1.tbl2.col1,db1.tbl2.col2,db1.tbl2.col3,db1.tbl2.col4", comments="")
    out = {"rows": n, "scripts": {}}
    res = {}
    for flavour in ("ref", "b200"):
        pair = H.ServerPair(flavour, os.path.join(work, flavour))
        pair.start()
        res[flavour] = {}
        for name, text in scripts(n, csv).items():
            reps = 1 if name == "load" else 5
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                o = pair.run_dsl(text)
                ts.append(time.perf_counter() - t0)
            res[flavour][name] = (o, ts)
        pair.stop()
    same = True
    for name in res["ref"]:
        o_ref, t_ref = res["ref"][name]
        o_gpu, t_gpu = res["b200"][name]
        same &= o_ref == o_gpu
        out["scripts"][name] = {"identical_output": o_ref == o_gpu, "output_bytes": len(o_ref),
                                "reference_ms": [round(1e3 * t, 2) for t in t_ref],
                                "b200_ms": [round(1e3 * t, 2) for t in t_gpu],
                                "reference_best_ms": round(1e3 * min(t_ref), 2),
                                "b200_best_ms": round(1e3 * min(t_gpu), 2)}
    out["all_outputs_identical"] = same
    out["note"] = ("wall clock of one client process per script; the first b200 run of a script that touches a "
                   "column for the first time includes its one-time H2D upload and, for the very first, CUDA "
                   "context creation; at 1 M rows both arms are dominated by process start, socket round trips "
                   "and parsing, not by the operators")
    print(json.dumps(out, indent=1))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
