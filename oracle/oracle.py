"""ctypes front-end for the CPU checker libraries.  TEST INFRASTRUCTURE ONLY.

Two libraries share one flat-array calling convention (see oracle/adb_oracle.c and
oracle/ref_harness.c):

* ``port()``      -> liboracle.so, the restatement (``orc_*`` symbols);
* ``reference()`` -> oracle/_ref/libref_O{0,2}.so, the UNMODIFIED reference operators
  compiled from /root/reference/src (``ref_*`` symbols); ``None`` if never built.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl
reference`` legs may import this module.  The engine package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I32P = C.POINTER(C.c_int32)
_U64P = C.POINTER(C.c_uint64)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_I32P)


def _bound(x):
    """None -> NULL pointer (bound absent, server.c:144-154); int -> int32*."""
    if x is None:
        return None, None
    box = C.c_int32(int(x))
    return C.pointer(box), box


class CpuOps:
    """The reference operator set over numpy int32 arrays."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.path, self.prefix, self.kind = path, prefix, kind
        self.lib = C.CDLL(path)
        f = self._fn
        f("select_scan", C.c_int64, [_I32P, C.c_int64, _I32P, _I32P, _I32P])
        f("select_result", C.c_int64, [_I32P, _I32P, C.c_int64, _I32P, _I32P, _I32P])
        f("select_sorted_index", C.c_int64,
          [_I32P, _U64P, C.c_int64, C.c_int32, C.c_int32, _I32P, _I32P])
        f("fetch", None, [_I32P, _I32P, C.c_int64, _I32P])
        f("sum", C.c_int64, [_I32P, C.c_int64])
        f("avg", C.c_double, [_I32P, C.c_int64])
        f("min", C.c_int32, [_I32P, C.c_int64])
        f("max", C.c_int32, [_I32P, C.c_int64])
        f("add", None, [_I32P, _I32P, C.c_int64, _I32P])
        f("sub", None, [_I32P, _I32P, C.c_int64, _I32P])
        join_args = [_I32P, _I32P, C.c_int64, _I32P, _I32P, C.c_int64,
                     C.POINTER(_I32P), C.POINTER(_I32P)]
        f("hash_join", C.c_int64, join_args)
        f("nested_loop_join", C.c_int64, join_args)
        f("free", None, [C.c_void_p])
        f("multimap_size", C.c_int32, [C.c_int32, C.c_int32])
        f("index_sort", None, [_I32P, C.c_int64, _I32P, _U64P])
        f("reorder", None, [_I32P, C.c_int64, _U64P])
        f("chain_select_fetch_sum", C.c_int64,
          [_I32P, _I32P, C.c_int64, _I32P, _I32P, C.POINTER(C.c_int64)])
        f("chain_select_fetch_sum_mt", C.c_int64,
          [_I32P, _I32P, C.c_int64, _I32P, _I32P, C.c_int32, C.POINTER(C.c_int64)])
        shared = [_I32P, C.c_int64, _I32P, _I32P, C.c_int32, C.POINTER(_I32P),
                  C.POINTER(C.c_int64)]
        if prefix == "ref_":
            shared += [C.c_int32, C.c_int32]
            f("sum_column", C.c_int64, [_I32P, C.c_int64])
        f("shared_select", None, shared)

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, "_" + name, fn)

    # ---- bulk load (restatement only: load_db is not part of the reference objects) ---
    def csv_parse(self, text: bytes, n_cols: int, skip_lines: int = 1) -> np.ndarray:
        """The ingest loop of load_db (db_manager.c:304-318) over the bytes of a CSV file:
        int32 array of shape (n_cols, rows)."""
        fn = self.lib.orc_csv_parse
        fn.restype = C.c_int64
        fn.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
        rows = fn(text, len(text), skip_lines, n_cols, None, 0)
        out = np.zeros((n_cols, max(rows, 1)), dtype=np.int32)
        fn(text, len(text), skip_lines, n_cols, out.ctypes.data, max(rows, 1))
        return out[:, :rows].copy()

    def print_i32(self, vals) -> bytes:
        """print of one INT result (query.c:245-304): the text, without a NUL."""
        vals = _i32(vals)
        fn = getattr(self.lib, self.prefix + "print_i32")
        fn.restype = C.c_int64
        fn.argtypes = [_I32P, C.c_int64, C.c_char_p, C.c_int64]
        buf = C.create_string_buffer(12 * max(vals.size, 1) + 16)
        n = fn(_p(vals), vals.size, buf, len(buf))
        return buf.raw[:n]

    # ---- selects -------------------------------------------------------------
    def select_scan(self, data, lo=None, hi=None) -> np.ndarray:
        data = _i32(data)
        out = np.empty(max(data.size, 1), dtype=np.int32)
        (plo, _k1), (phi, _k2) = _bound(lo), _bound(hi)
        h = self._select_scan(_p(data), data.size, plo, phi, _p(out))
        return out[:h].copy()

    def select_result(self, val, pos, lo=None, hi=None) -> np.ndarray:
        val, pos = _i32(val), _i32(pos)
        out = np.empty(max(val.size, 1), dtype=np.int32)
        (plo, _k1), (phi, _k2) = _bound(lo), _bound(hi)
        h = self._select_result(_p(val), _p(pos), val.size, plo, phi, _p(out))
        return out[:h].copy()

    def select_sorted_index(self, values, positions, lo: int, hi: int):
        """Returns (positions, undefined_flag)."""
        values = _i32(values)
        positions = np.ascontiguousarray(positions, dtype=np.uint64)
        out = np.empty(max(values.size, 1), dtype=np.int32)
        undef = C.c_int32(0)
        h = self._select_sorted_index(_p(values), positions.ctypes.data_as(_U64P), values.size,
                                      int(lo), int(hi), _p(out), C.byref(undef))
        return out[:h].copy(), bool(undef.value)

    def shared_select(self, data, lows, highs, col_min=None, col_max=None):
        data, lows, highs = _i32(data), _i32(lows), _i32(highs)
        q = lows.size
        bufs = [np.empty(max(data.size, 1), dtype=np.int32) for _ in range(q)]
        ptrs = (_I32P * q)(*[_p(b) for b in bufs])
        counts = (C.c_int64 * q)()
        args = [_p(data), data.size, _p(lows), _p(highs), q, ptrs, counts]
        if self.prefix == "ref_":
            cmin = int(data.min()) if col_min is None else col_min
            cmax = int(data.max()) if col_max is None else col_max
            args += [cmin, cmax]
        self._shared_select(*args)
        return [bufs[i][:counts[i]].copy() for i in range(q)]

    # ---- fetch / aggregates / arithmetic ----------------------------------------
    def fetch(self, data, pos) -> np.ndarray:
        data, pos = _i32(data), _i32(pos)
        out = np.empty(max(pos.size, 1), dtype=np.int32)
        self._fetch(_p(data), _p(pos), pos.size, _p(out))
        return out[:pos.size].copy()

    def sum(self, v) -> int:
        v = _i32(v)
        return int(self._sum(_p(v), v.size))

    def sum_column(self, v) -> int:
        """sum over a whole Column (query.c:336-341); same loop as over a Result."""
        v = _i32(v)
        fn = self._sum_column if self.prefix == "ref_" else self._sum
        return int(fn(_p(v), v.size))

    def avg(self, v) -> float:
        v = _i32(v)
        return float(self._avg(_p(v), v.size))

    def min(self, v) -> int:
        v = _i32(v)
        return int(self._min(_p(v), v.size))

    def max(self, v) -> int:
        v = _i32(v)
        return int(self._max(_p(v), v.size))

    def add(self, a, b) -> np.ndarray:
        a, b = _i32(a), _i32(b)
        out = np.empty(max(a.size, 1), dtype=np.int32)
        self._add(_p(a), _p(b), a.size, _p(out))
        return out[:a.size].copy()

    def sub(self, a, b) -> np.ndarray:
        a, b = _i32(a), _i32(b)
        out = np.empty(max(a.size, 1), dtype=np.int32)
        self._sub(_p(a), _p(b), a.size, _p(out))
        return out[:a.size].copy()

    # ---- joins -----------------------------------------------------------------
    def _join(self, fn, v1, p1, v2, p2):
        v1, p1, v2, p2 = _i32(v1), _i32(p1), _i32(v2), _i32(p2)
        o1, o2 = _I32P(), _I32P()
        m = fn(_p(v1), _p(p1), v1.size, _p(v2), _p(p2), v2.size, C.byref(o1), C.byref(o2))
        a = np.ctypeslib.as_array(o1, shape=(m,)).copy() if m else np.empty(0, np.int32)
        b = np.ctypeslib.as_array(o2, shape=(m,)).copy() if m else np.empty(0, np.int32)
        self._free(C.cast(o1, C.c_void_p))
        self._free(C.cast(o2, C.c_void_p))
        return a, b

    def hash_join(self, v1, p1, v2, p2):
        return self._join(self._hash_join, v1, p1, v2, p2)

    def nested_loop_join(self, v1, p1, v2, p2):
        return self._join(self._nested_loop_join, v1, p1, v2, p2)

    def multimap_size(self, n: int, exact: bool = True) -> int:
        return int(self._multimap_size(int(n), int(exact)))

    # ---- index build -------------------------------------------------------------
    def index_sort(self, data):
        data = _i32(data)
        values = np.empty(max(data.size, 1), dtype=np.int32)
        positions = np.empty(max(data.size, 1), dtype=np.uint64)
        self._index_sort(_p(data), data.size, _p(values), positions.ctypes.data_as(_U64P))
        return values[:data.size].copy(), positions[:data.size].copy()

    def reorder(self, data, sorted_positions) -> np.ndarray:
        out = _i32(data).copy()
        sp = np.ascontiguousarray(sorted_positions, dtype=np.uint64)
        self._reorder(_p(out), out.size, sp.ctypes.data_as(_U64P))
        return out

    # ---- the timed chain -----------------------------------------------------------
    def chain_select_fetch_sum(self, sel, fet, lo=None, hi=None, threads: int = 1):
        """Returns (sum, hits) of select(sel, lo, hi) -> fetch(fet) -> sum."""
        sel, fet = _i32(sel), _i32(fet)
        (plo, _k1), (phi, _k2) = _bound(lo), _bound(hi)
        hits = C.c_int64(0)
        if threads <= 1:
            s = self._chain_select_fetch_sum(_p(sel), _p(fet), sel.size, plo, phi, C.byref(hits))
        else:
            s = self._chain_select_fetch_sum_mt(_p(sel), _p(fet), sel.size, plo, phi,
                                                int(threads), C.byref(hits))
        return int(s), int(hits.value)


def synth_uniform(n: int, seed: int, first_row: int, lo: int, span: int, threads: int = 1) -> np.ndarray:
    """The generator of analytical-database_b200/synth.py on `threads` host threads (C, in
    liboracle.so): input regeneration for the CPU arm of bench.py at full table size."""
    lib = port().lib
    fn = lib.orc_synth_uniform_mt
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_int32, C.c_uint32, C.c_int]
    out = np.empty(n, dtype=np.int32)
    fn(out.ctypes.data_as(C.c_void_p), n, seed, first_row, lo, span, threads)
    return out


_cache: dict = {}


def port() -> CpuOps:
    """The restatement (always available; built on demand)."""
    if "port" not in _cache:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _cache["port"] = CpuOps(path, "orc_", "port")
    return _cache["port"]


def reference(opt: str = "O2"):
    """The unmodified reference operators, or None when oracle/_ref was never built."""
    key = "ref_" + opt
    if key not in _cache:
        path = os.path.join(_HERE, "_ref", f"libref_{opt}.so")
        _cache[key] = CpuOps(path, "ref_", "reference") if os.path.exists(path) else None
    return _cache[key]
