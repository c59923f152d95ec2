"""Fixtures for the bulk-load path (adb_csv_index / adb_csv_parse, oracle orc_csv_parse).

Writes a handful of CSV files under tests/golden/load/ -- one plain, the rest exercising what
the reference's ingest loop (/root/reference/src/db_manager.c:304-318: fgets, strsep at ',',
atoi per token, insert_row) does with untidy input -- and asks the UNMODIFIED reference
server (oracle/_ref/dropin/server_ref, built by oracle/Makefile from /root/reference/src) to
load each one and print every column.  What it printed is committed next to the CSV as
<name>.cols.json: the ground truth the oracle's restatement is pinned against
(tests/test_csv_load.py).  Needs /root/reference; run from the repo root:

    python tests/golden/make_golden_load.py
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from dsl_harness import ServerPair  # noqa: E402

OUT = os.path.join(HERE, "load")


def cases():
    rng = np.random.default_rng(42)
    plain = ["%d,%d,%d,%d" % tuple(r) for r in
             np.stack([rng.integers(-500, 500, 400), rng.integers(0, 100, 400),
                       rng.integers(-10**6, 10**6, 400), rng.integers(0, 10, 400)], 1)]
    yield "plain", 4, "\n".join(plain) + "\n"
    # no newline after the last row: fgets still returns it
    yield "no_final_newline", 3, "1,2,3\n4,5,6\n-7,-8,-9"
    messy = [
        "1,2,3",
        " 12,\t-5,+7",                 # leading whitespace, explicit sign
        "12abc,abc,",                  # digits then junk; junk; empty token
        "--5,- 5,007",                 # double sign, sign then space, leading zeros
        "4 ,5 5,6\t",                  # trailing junk after the digits
        "10,20,30,40,50",              # more fields than columns: ignored
        "7,8",                         # short row: the third column keeps the previous row's 30
        "9",                           # shorter still: 8 and 30 are kept
        "",                            # empty line: token "\n" -> 0, the rest kept
        "99999999999,2147483647,2147483648",          # (int)(long): truncation
        "9223372036854775807,9223372036854775808,-9223372036854775809",   # strtol saturates
        "-2147483648,-2147483649,18446744073709551616",
        "3,\v4,\f5",                   # the other isspace() characters
        "1.5,2e3,0x10",                # atoi stops at '.', 'e', 'x'
        "5,6,7",
    ]
    yield "messy", 3, "\n".join(messy) + "\n"
    yield "crlf", 2, "1,2\r\n3,4\r\n-5,-6\r\n"
    yield "one_column", 1, "\n".join(str(int(x)) for x in rng.integers(-99, 99, 50)) + "\n"


def reference_columns(work, name, n_cols, csv_path):
    """create + load + shutdown in one client session, restart, query -- the protocol of the
    reference's own suite (infra_scripts/test_milestone.sh; a fresh server per case keeps
    the cases independent)."""
    shutil.rmtree(os.path.join(work, "database"), ignore_errors=True)
    pair = ServerPair("ref", work)
    cols = [f"c{i}" for i in range(n_cols)]
    dsl = ['create(db,"db1")', f'create(tbl,"{name}",db1,{n_cols})']
    dsl += [f'create(col,"{c}",db1.{name})' for c in cols]
    dsl += [f'load("{csv_path}")', "shutdown"]
    try:
        pair.start()
        pair.run_dsl("\n".join(dsl) + "\n")
        pair.start()
        got = []
        for c in cols:
            out = pair.run_dsl(f"s=select(db1.{name}.{c},null,null)\nf=fetch(db1.{name}.{c},s)\nprint(f)\n")
            try:
                got.append([int(x) for x in out.split()])
            except ValueError:
                raise SystemExit(f"{name}.{c}: the reference server answered {out!r}")
    finally:
        pair.stop()
    return got


def main():
    os.makedirs(OUT, exist_ok=True)
    if not ServerPair.available("ref"):
        raise SystemExit("oracle/_ref/dropin/server_ref is not built (make -C oracle)")
    with tempfile.TemporaryDirectory() as work:
        for name, n_cols, body in cases():
            header = ",".join(f"db1.{name}.c{i}" for i in range(n_cols))
            path = os.path.join(OUT, name + ".csv")
            with open(path, "w", newline="") as f:
                f.write(header + "\n" + body)
            cols = reference_columns(work, name, n_cols, path)
            with open(os.path.join(OUT, name + ".cols.json"), "w") as f:
                json.dump({"n_cols": n_cols, "rows": len(cols[0]), "columns": cols}, f)
            print(name, n_cols, "cols", [len(c) for c in cols], "rows")


if __name__ == "__main__":
    main()
