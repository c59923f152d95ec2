"""Result transport end to end: a long `print` travels from the engine
through the reply path of the reference's server and client.  The unchanged path copies the
reply into a stack VLA (server.c:521-524) and reads it with a single recv into another
(client.c:126-133); patches/apply_reply_patch.py removes both (payload sent / received in a
loop, heap buffer).  oracle/Makefile builds the drop-in with the patched copies as
server_chunked / client_chunked.

  * multi-MB prints arrive complete and byte-exact (the text is formatted on the device);
  * on replies the unpatched pair survives, the patched pair prints the same bytes as the
    unmodified reference pair."""
import os
import tempfile

import numpy as np
import pytest

import dsl_harness as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (H.ServerPair.available("chunked") and H.ServerPair.available("ref")),
                                 reason="oracle/_ref/dropin not built")]


def write_table(path, cols):
    n = cols[0].size
    with open(path, "w") as f:
        f.write(",".join(f"db1.tbl1.col{j + 1}" for j in range(len(cols))) + "\n")
        np.savetxt(f, np.stack(cols, axis=1), fmt="%d", delimiter=",")
    return n


def script(csv, ncols, queries):
    head = ['create(db,"db1")', f'create(tbl,"tbl1",db1,{ncols})'] + \
           [f'create(col,"col{j + 1}",db1.tbl1)' for j in range(ncols)] + [f'load("{csv}")']
    return "\n".join(head + queries) + "\n"


@pytest.mark.parametrize("gpus", ["1", "2"])
@pytest.mark.parametrize("rows", [300_000, 4_000_000])
def test_long_print_travels_end_to_end(rows, gpus):
    rng = np.random.default_rng(rows)
    c1 = rng.integers(-2**31, 2**31 - 1, rows, dtype=np.int64).astype(np.int32)
    c2 = rng.integers(0, 1000, rows).astype(np.int32)
    with tempfile.TemporaryDirectory(prefix="adb_chunked_") as w:
        csv = os.path.join(w, "t.csv")
        write_table(csv, [c1, c2])
        pair = H.ServerPair("chunked", w, env={"ADB_GPUS": gpus})
        try:
            out = pair.run_dsl(script(csv, 2, ["s1=select(db1.tbl1.col2,100,900)", "f1=fetch(db1.tbl1.col1,s1)",
                                               "print(f1)", "a1=sum(f1)", "print(a1)"]), timeout=600)
        finally:
            pair.stop()
    sel = (c2 >= 100) & (c2 < 900)
    exp = "\n".join(map(str, c1[sel].tolist())) + "\n" + str(int(c1[sel].astype(np.int64).sum())) + "\n"
    # ~85 % of the rows, 10-12 bytes each: 3 MB and 40 MB of text
    assert len(out) == len(exp)
    assert out == exp


def test_patched_pair_prints_what_the_reference_pair_prints():
    """A reply both paths can carry (a few KB): same bytes from the unmodified reference server +
    client and from the patched drop-in pair."""
    rows = 3000
    rng = np.random.default_rng(7)
    c1 = rng.integers(-10**6, 10**6, rows).astype(np.int32)
    c2 = rng.integers(0, 100, rows).astype(np.int32)
    queries = ["s1=select(db1.tbl1.col2,10,60)", "f1=fetch(db1.tbl1.col1,s1)", "print(f1)", "a1=avg(f1)",
               "print(a1)", "m1=min(f1)", "m2=max(f1)", "print(m1,m2)"]
    outs = {}
    for flavour in ("ref", "chunked"):
        with tempfile.TemporaryDirectory(prefix=f"adb_{flavour}_") as w:
            csv = os.path.join(w, "t.csv")
            write_table(csv, [c1, c2])
            pair = H.ServerPair(flavour, w)
            try:
                outs[flavour] = pair.run_dsl(script(csv, 2, queries))
            finally:
                pair.stop()
    assert outs["chunked"] == outs["ref"] and len(outs["ref"]) > 5000


@pytest.mark.skipif(not os.environ.get("ADB_TEST_BIG_PRINT"), reason="set ADB_TEST_BIG_PRINT=1: 50 M rows, ~1 GB of text, minutes")
def test_fifty_million_row_print_travels_end_to_end():
    """a 50 M-row print through the patched reply path (550 MB of text in one
    reply, formatted on the device, sent and received in pieces).  Slow (the reference's load_db
    parses 50 M CSV lines with atoi, and the client's output is captured), so opt-in; the result of
    the round's run is recorded in profiles/r02_print_50M.md."""
    import time
    rows = 50_000_000
    rng = np.random.default_rng(50)
    c1 = rng.integers(-2**31, 2**31 - 1, rows, dtype=np.int64).astype(np.int32)
    with tempfile.TemporaryDirectory(prefix="adb_chunked_big_") as w:
        csv = os.path.join(w, "t.csv")
        with open(csv, "w") as f:
            f.write("db1.tbl1.col1\n")
            for b in range(0, rows, 1 << 22):
                f.write("\n".join(map(str, c1[b:b + (1 << 22)].tolist())) + "\n")
        pair = H.ServerPair("chunked", w)
        try:
            t0 = time.time()
            pair.run_dsl(script(csv, 1, []), timeout=1800)
            t_load = time.time() - t0
            t0 = time.time()
            out = pair.run_dsl("s1=select(db1.tbl1.col1,null,null)\nf1=fetch(db1.tbl1.col1,s1)\nprint(f1)\n", timeout=1800)
            t_print = time.time() - t0
        finally:
            pair.stop()
    got = np.fromstring(out, sep="\n", dtype=np.int64)
    assert got.size == rows and np.array_equal(got.astype(np.int32), c1)
    print(f"50M-row print: load {t_load:.1f} s, select+fetch+print round trip {t_print:.1f} s, reply {len(out)} bytes")
