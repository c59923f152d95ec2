"""Oracle-backed stand-in for sharded.EngineOps: the same local-operator interface on CPU
torch tensors, so the partitioning / exchange logic of analytical-database_b200/sharded.py can
run under gloo without a GPU.  Test infrastructure (it imports the oracle)."""
import numpy as np
import torch

from oracle import oracle


def route_dest(keys: np.ndarray, parts: int) -> np.ndarray:
    """Destination of adb_route_pairs: top log2(parts) bits of key * 0x85EBCA6B (mod 2^32)."""
    bits = parts.bit_length() - 1
    if bits == 0:
        return np.zeros(keys.size, dtype=np.int64)
    h = (keys.astype(np.int64).astype(np.uint64) * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    return (h >> np.uint64(32 - bits)).astype(np.int64)


class OracleOps:
    device = torch.device("cpu")

    def __init__(self):
        self.o = oracle.port()

    @staticmethod
    def _t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32))

    def select_scan(self, col, lo, hi):
        return self._t(self.o.select_scan(col.numpy(), lo, hi))

    def fetch(self, col, pos):
        return self._t(self.o.fetch(col.numpy(), pos.numpy()))

    def aggregate_packed(self, vals):
        v = vals.numpy()
        mx = self.o.max(v) if v.size else -2**31
        mn = self.o.min(v) if v.size else 2**31 - 1
        return (torch.tensor([self.o.sum(v), v.size], dtype=torch.int64),
                torch.tensor([mx, ~mn], dtype=torch.int32))

    def ewise(self, a, b, subtract):
        return self._t((self.o.sub if subtract else self.o.add)(a.numpy(), b.numpy()))

    def shared_select(self, col, lows, highs):
        return [self._t(x) for x in self.o.shared_select(col.numpy(), lows, highs)]

    def index_build(self, col, with_btree=True):
        a = col.numpy()
        order = np.argsort(a, kind="stable")
        return self._t(a[order]), self._t(order), None

    def select_index(self, index, lo, hi, use_btree):
        vals, poss = index[0].numpy(), index[1].numpy()
        l = 0 if lo is None else int(np.searchsorted(vals, lo, "left"))
        h = vals.size if hi is None else int(np.searchsorted(vals, hi, "left"))
        return self._t(poss[l:max(l, h)])

    def route_pairs(self, val, pos, parts):
        d = route_dest(val.numpy(), parts)
        order = np.argsort(d, kind="stable")
        counts = np.bincount(d, minlength=parts).tolist()
        return self._t(val.numpy()[order]), self._t(pos.numpy()[order]), counts

    def hash_join(self, v1, p1, v2, p2):
        a, b = self.o.hash_join(v1.numpy(), p1.numpy(), v2.numpy(), p2.numpy()) if v1.numel() and v2.numel() \
            else (np.empty(0, np.int32), np.empty(0, np.int32))
        return self._t(a), self._t(b)
