"""Bulk load: CSV text -> int32 columns.

CPU (-m "not gpu"): the oracle's restatement of load_db's ingest loop
(/root/reference/src/db_manager.c:304-318) against what the UNMODIFIED reference server
printed after loading the same files (tests/golden/load/*.cols.json, made by
tests/golden/make_golden_load.py).
GPU (-m gpu): adb_csv_index + adb_csv_parse through the C-ABI against those fixtures and,
on seeded random text, against the oracle -- tidy tables, untidy tokens, short rows, lines
at the fgets limit, a file large enough for the staged (pinned, multi-lane) upload."""
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LOAD = os.path.join(HERE, "golden", "load")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(LOAD, "*.csv")))


def fixture(name):
    with open(os.path.join(LOAD, name + ".csv"), "rb") as f:
        text = f.read()
    with open(os.path.join(LOAD, name + ".cols.json")) as f:
        exp = json.load(f)
    return text, exp["n_cols"], np.array(exp["columns"], dtype=np.int64).astype(np.int32).reshape(exp["n_cols"], -1)


def test_fixtures_exist():
    assert {"plain", "messy", "crlf", "no_final_newline", "one_column"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_parse_equals_reference_server(port, name):
    text, n_cols, exp = fixture(name)
    got = port.csv_parse(text, n_cols)
    assert got.shape == exp.shape and np.array_equal(got, exp)


def test_oracle_parse_edges(port):
    assert port.csv_parse(b"", 2).shape == (2, 0)
    assert port.csv_parse(b"h\n", 2).shape == (2, 0)                 # header only
    assert port.csv_parse(b"h\n\n", 2).tolist() == [[0], [0]]        # one empty line is a row
    assert port.csv_parse(b"1,2\n3,4", 2, skip_lines=0).tolist() == [[1, 3], [2, 4]]
    long_line = b"h\n" + b"1," * 600 + b"\n"                         # 1201 bytes: fgets splits it
    assert port.csv_parse(long_line, 1).shape[1] == 2


TOKENS = ["0", "7", "-7", "+7", " 42", "\t-3", "12abc", "abc", "", "--5", "- 5", "007", "4 ", "2147483647",
          "2147483648", "-2147483648", "-2147483649", "99999999999", "9223372036854775807",
          "9223372036854775808", "-9223372036854775809", "18446744073709551616", "1.5", "2e3", "0x10",
          "\v4", "\f5", "\r6", "5\r"]


def random_text(rng, rows, n_cols, tidy):
    lines = ["db1.t." + ",db1.t.".join(f"c{i}" for i in range(n_cols))]
    for r in range(rows):
        if tidy:
            k = n_cols
            toks = [str(int(x)) for x in rng.integers(-2**31, 2**31 - 1, k)]
        else:
            k = n_cols if r == 0 else int(rng.integers(0, n_cols + 3))     # short / long rows after the first
            toks = [TOKENS[int(i)] for i in rng.integers(0, len(TOKENS), k)]
        lines.append(",".join(toks))
    return ("\n".join(lines) + ("\n" if rng.integers(0, 2) else "")).encode()


def engine_columns(eng, text, n_cols, skip=1):
    cols, rows = eng.csv_load(text, n_cols, skip)
    out = np.stack([c.to_host(rows) for c in cols]) if rows else np.zeros((n_cols, 0), np.int32)
    for c in cols:
        c.free()
    return out


@pytest.fixture(scope="module")
def eng():
    from analytical_database_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_parse_equals_reference_server(eng, name):
    text, n_cols, exp = fixture(name)
    got = engine_columns(eng, text, n_cols)
    assert got.shape == exp.shape and np.array_equal(got, exp)


@pytest.mark.gpu
@pytest.mark.parametrize("tidy", [True, False])
@pytest.mark.parametrize("rows,n_cols", [(0, 1), (1, 1), (3, 4), (255, 2), (256, 3), (257, 7), (5000, 4), (70001, 5)])
def test_engine_parse_equals_oracle(eng, port, rng, rows, n_cols, tidy):
    text = random_text(rng, rows, n_cols, tidy)
    exp = port.csv_parse(text, n_cols)
    got = engine_columns(eng, text, n_cols)
    assert got.shape == exp.shape and np.array_equal(got, exp)


@pytest.mark.gpu
def test_engine_parse_edges(eng, port):
    from analytical_database_b200 import EngineError
    for text, n_cols, skip in [(b"", 2, 1), (b"h\n", 2, 1), (b"h\n\n", 2, 1), (b"1,2\n3,4", 2, 0),
                               (b"h", 1, 1), (b"\n\n\n", 3, 0), (b"5", 1, 0)]:
        assert np.array_equal(engine_columns(eng, text, n_cols, skip), port.csv_parse(text, n_cols, skip)), text
    ok = b"h\n" + b"1," * 510 + b"2\n"                                 # 1022 bytes + '\n' = 1023: the limit
    assert np.array_equal(engine_columns(eng, ok, 3), port.csv_parse(ok, 3))
    with pytest.raises(EngineError, match="1023"):
        engine_columns(eng, b"h\n" + b"1," * 600 + b"\n", 1)
    with pytest.raises(EngineError):
        eng.csv_load(b"1\n", 255, 0)


@pytest.mark.gpu
def test_engine_parse_large_file_through_the_staged_upload(eng, port):
    """~80 MB of text (> the 64 MB staging threshold): 4 columns x 3 M rows."""
    rng = np.random.default_rng(7)
    rows = 3_000_000
    cols = [rng.integers(-2**31, 2**31 - 1, rows).astype(np.int32) for _ in range(2)]
    cols += [rng.integers(0, 100, rows).astype(np.int32), np.arange(rows, dtype=np.int32)]
    import pandas as pd
    text = ("a,b,c,d\n" + pd.DataFrame(dict(zip("abcd", cols))).to_csv(index=False, header=False)).encode()
    assert len(text) > 64 << 20
    got = engine_columns(eng, text, 4)
    assert got.shape == (4, rows)
    for g, c in zip(got, cols):
        assert np.array_equal(g, c)
    # and the staged download path brings a large column back intact (to_host above used it
    # only for 12 MB pieces): round-trip 80 MB
    a = np.frombuffer(text, dtype=np.uint8)
    d = eng.upload(a)
    assert np.array_equal(d.to_host(a.size, np.uint8), a)
    d.free()
