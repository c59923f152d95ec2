"""ctypes view of include/adb_query_api.h: the reference's operator types and the thirteen
query.h functions as libadb_query.so (host/query_shim.c) exports them.  Test plumbing."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "analytical-database_b200", "libadb_query.so")
INT, LONG, FLOAT, DOUBLE = 0, 1, 2, 3
OK, ERROR = 0, 1
RESULT, COLUMN = 0, 1


class ColumnIndex(C.Structure):
    _fields_ = [("values", C.POINTER(C.c_int)), ("positions", C.POINTER(C.c_size_t))]


class Column(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("data", C.POINTER(C.c_int)), ("fd", C.c_int),
                ("row_count", C.c_size_t), ("sorted", C.c_bool), ("clustered", C.c_bool),
                ("has_index", C.c_bool), ("index", C.POINTER(ColumnIndex)),
                ("btree_node", C.c_void_p), ("histogram", C.c_void_p), ("max", C.c_int),
                ("min", C.c_int)]


class Status(C.Structure):
    _fields_ = [("code", C.c_int), ("error_message", C.c_char_p)]


class Result(C.Structure):
    _fields_ = [("num_tuples", C.c_size_t), ("data_type", C.c_int), ("payload", C.c_void_p)]


class GeneralizedColumnPointer(C.Union):
    _fields_ = [("result", C.POINTER(Result)), ("column", C.POINTER(Column))]


class GeneralizedColumn(C.Structure):
    _fields_ = [("column_type", C.c_int), ("column_pointer", GeneralizedColumnPointer)]


class SelectOperator(C.Structure):
    _fields_ = [("select_type", C.c_int), ("handle", C.c_char * 64), ("low", C.c_int),
                ("high", C.c_int), ("has_low", C.c_int), ("has_high", C.c_int),
                ("db", C.c_void_p), ("table", C.c_void_p), ("column", C.POINTER(Column)),
                ("col_result", C.POINTER(Result)), ("pos_result", C.POINTER(Result)),
                ("comparator", C.c_void_p)]


RP, RPP, IP = C.POINTER(Result), C.POINTER(C.POINTER(Result)), C.POINTER(C.c_int)
SP = C.POINTER(Status)
OPERATORS = {
    "select_result": (RP, [RP, RP, IP, IP, SP]),
    "select_column": (RP, [C.POINTER(Column), IP, IP, SP]),
    "fetch_column": (RP, [C.POINTER(Column), RP, SP]),
    "print": (C.c_void_p, [RPP, C.c_int, SP]),
    "average": (RP, [RP, SP]),
    "sum": (RP, [C.POINTER(GeneralizedColumn), SP]),
    "add": (RP, [RP, RP, SP]),
    "sub": (RP, [RP, RP, SP]),
    "min": (RP, [RP, SP]),
    "max": (RP, [RP, SP]),
    "shared_select": (RPP, [C.POINTER(SelectOperator), C.c_int, C.POINTER(Column), SP]),
    "nested_loop_join": (RPP, [RP, RP, RP, RP, SP]),
    "hash_join": (RPP, [RP, RP, RP, RP, SP]),
    "log_result": (None, [RP]),
    "should_use_index": (C.c_bool, [C.POINTER(Column), C.c_int, C.c_int]),
}
HOOKS = {
    "adb_host_init": (C.c_int, [C.c_int]),
    "adb_host_init_multi": (C.c_int, [C.c_int]),
    "adb_host_gpus": (C.c_int, []),
    "adb_host_column_adopt_shards": (C.c_int, [C.POINTER(Column), C.POINTER(C.c_void_p), C.c_size_t]),
    "adb_host_shutdown": (None, []),
    "adb_host_column_upload": (C.c_int, [C.POINTER(Column)]),
    "adb_host_column_adopt": (C.c_int, [C.POINTER(Column), C.c_void_p]),
    "adb_host_column_invalidate": (None, [C.POINTER(Column)]),
    "adb_host_index_build": (C.c_int, [C.POINTER(C.POINTER(Column)), C.c_int, C.c_int]),
    "adb_host_relational_update": (C.c_int, [C.POINTER(C.POINTER(Column)), C.c_int, C.c_int, RP, C.c_int]),
    "adb_host_relational_delete": (C.c_int, [C.POINTER(C.POINTER(Column)), C.c_int, RP]),
    "adb_host_column_histogram": (C.c_int, [C.POINTER(Column), C.c_int, C.POINTER(C.c_ulong)]),
    "adb_host_result_release": (None, [RP]),
    "adb_host_payload_freed": (None, [C.c_void_p]),
    "adb_host_results_drop": (None, [RPP, C.c_int]),
    "adb_host_result_to_host": (C.c_int, [RP, C.c_void_p]),
    "adb_host_last_error": (C.c_char_p, []),
    "adb_host_live_device_results": (C.c_long, []),
    "adb_host_profile_dump": (None, []),
}


def load():
    lib = C.CDLL(LIB)
    for name, (res, args) in {**OPERATORS, **HOOKS}.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


class Api:
    """The operator API over numpy arrays, the way the dispatcher drives it."""

    def __init__(self):
        self.lib = load()
        self._keep = []

    def column(self, data, index=None, sorted_=False, clustered=False) -> Column:
        """index = (values int32[n], positions uint64[n]) as src/index.c builds them."""
        data = np.ascontiguousarray(data, dtype=np.int32)
        col = Column()
        col.name = b"col"
        col.data = data.ctypes.data_as(C.POINTER(C.c_int))
        col.row_count = data.size
        col.sorted, col.clustered, col.has_index = sorted_, clustered, index is not None
        if data.size:
            col.min, col.max = int(data.min()), int(data.max())
        self._keep.append(data)
        if index is not None:
            v = np.ascontiguousarray(index[0], dtype=np.int32)
            p = np.ascontiguousarray(index[1], dtype=np.uint64)
            ix = ColumnIndex(v.ctypes.data_as(C.POINTER(C.c_int)), p.ctypes.data_as(C.POINTER(C.c_size_t)))
            col.index = C.pointer(ix)
            self._keep += [v, p, ix]
        return col

    def table(self, arrays, flags=None):
        """A table's columns in declaration order (writable host arrays: a clustered index build
        permutes sibling data in place).  flags[j] = (sorted, clustered) for an indexed column."""
        cols = []
        for j, a in enumerate(arrays):
            a = np.array(a, dtype=np.int32, copy=True)
            col = Column()
            col.name = b"col%d" % (j + 1)
            col.data = a.ctypes.data_as(C.POINTER(C.c_int))
            col.row_count = a.size
            if a.size:
                col.min, col.max = int(a.min()), int(a.max())
            if flags and flags.get(j):
                col.sorted, col.clustered = flags[j]
                col.has_index = True
            self._keep.append(a)
            cols.append((col, a))
        return cols

    def build_index(self, cols, which):
        arr = (C.POINTER(Column) * len(cols))(*[C.pointer(c) for c, _ in cols])
        if self.lib.adb_host_index_build(arr, len(cols), which) != 0:
            raise RuntimeError("adb_host_index_build: " + self.lib.adb_host_last_error().decode())
        col = cols[which][0]
        n = col.row_count
        ix = col.index.contents
        return (np.ctypeslib.as_array(ix.values, shape=(n,)).copy() if n else np.zeros(0, np.int32),
                np.ctypeslib.as_array(ix.positions, shape=(n,)).copy() if n else np.zeros(0, np.uint64))

    def host_result(self, values) -> Result:
        """A Result whose payload is an ordinary host int array (as the reference builds)."""
        values = np.ascontiguousarray(values, dtype=np.int32)
        self._keep.append(values)
        return Result(values.size, INT, values.ctypes.data_as(C.c_void_p).value)

    @staticmethod
    def _b(x):
        return None if x is None else C.pointer(C.c_int(int(x)))

    def check(self, res, st, what):
        if st.code != OK or not res:
            raise RuntimeError(f"{what}: {self.lib.adb_host_last_error().decode()}")
        return res

    def select_column(self, col, lo=None, hi=None):
        st = Status(99, None)
        return self.check(self.lib.select_column(C.byref(col), self._b(lo), self._b(hi), C.byref(st)), st, "select_column")

    def select_result(self, val, pos, lo=None, hi=None):
        st = Status(99, None)
        return self.check(self.lib.select_result(val, pos, self._b(lo), self._b(hi), C.byref(st)), st, "select_result")

    def fetch_column(self, col, pos):
        st = Status(99, None)
        return self.check(self.lib.fetch_column(C.byref(col), pos, C.byref(st)), st, "fetch_column")

    def unary(self, name, r):
        st = Status(99, None)
        return self.check(getattr(self.lib, name)(r, C.byref(st)), st, name)

    def sum_result(self, r):
        g = GeneralizedColumn(RESULT)
        g.column_pointer.result = r
        st = Status(99, None)
        return self.check(self.lib.sum(C.byref(g), C.byref(st)), st, "sum")

    def sum_column(self, col):
        g = GeneralizedColumn(COLUMN)
        g.column_pointer.column = C.pointer(col)
        st = Status(99, None)
        return self.check(self.lib.sum(C.byref(g), C.byref(st)), st, "sum")

    def binary(self, name, a, b):
        st = Status(99, None)
        return self.check(getattr(self.lib, name)(a, b, C.byref(st)), st, name)

    def shared_select(self, col, lows, highs):
        q = len(lows)
        ops = (SelectOperator * q)()
        for i in range(q):
            ops[i].low, ops[i].high, ops[i].has_low, ops[i].has_high = int(lows[i]), int(highs[i]), 1, 1
        st = Status(99, None)
        res = self.check(self.lib.shared_select(ops, q, C.byref(col), C.byref(st)), st, "shared_select")
        out = [C.pointer(res[i].contents) for i in range(q)]    # detach from the array about to be freed
        _libc.free(C.cast(res, C.c_void_p))
        return out

    def join(self, name, v1, p1, v2, p2):
        st = Status(99, None)
        res = self.check(getattr(self.lib, name)(v1, p1, v2, p2, C.byref(st)), st, name)
        out = (C.pointer(res[0].contents), C.pointer(res[1].contents))
        _libc.free(C.cast(res, C.c_void_p))      # src/server.c:432
        return out

    def print(self, *results) -> str:
        arr = (RP * len(results))(*results)
        st = Status(99, None)
        p = self.lib.print(arr, len(results), C.byref(st))
        if st.code != OK or not p:
            raise RuntimeError(f"print: {self.lib.adb_host_last_error().decode()}")
        s = C.string_at(p).decode()
        _libc.free(p)                             # src/server.c:539-541
        return s

    # ---- what a test reads back ----------------------------------------------------
    def tuples(self, r) -> np.ndarray:
        r = r.contents if isinstance(r, RP) else r
        dt = {INT: np.int32, LONG: np.int64, FLOAT: np.float32, DOUBLE: np.float64}[r.data_type]
        out = np.empty(r.num_tuples, dtype=dt)
        assert self.lib.adb_host_result_to_host(C.byref(r), out.ctypes.data_as(C.c_void_p)) == 0
        return out

    def drop(self, r):
        """What update_result / free_client_context do (src/client_context.c:31-45,76-90),
        with the two-line patch applied."""
        self.lib.adb_host_result_release(r)
        _libc.free(r.contents.payload)
        _libc.free(C.cast(r, C.c_void_p))
