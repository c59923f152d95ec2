"""world_size-2 gloo tests of the N > 1 host logic (sharded.py): row-range partitioning, count
offsets, aggregate all-reduce packing, batched-count exchange, and the hash-partitioned
all-to-all join, with the CPU oracle standing in for the local CUDA operators.  The result of
the 2-shard run must equal the oracle's 1-shard result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N = 40_007


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _table(seed=42):
    rng = np.random.default_rng(seed)
    return {"c1": rng.integers(-N // 2, N // 2, N).astype(np.int32),
            "c2": rng.integers(2**31 - 10000, 2**31 - 1, N, dtype=np.int64).astype(np.int32),
            "k": rng.integers(1, 5000, N).astype(np.int32)}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from analytical_database_b200.sharded import ShardedTable, shard_range
    from sharded_cpu_ops import OracleOps
    tab = _table()
    b, e = shard_range(N, rank, world)
    t = ShardedTable(OracleOps(), {k: torch.from_numpy(v[b:e].copy()) for k, v in tab.items()}, N, dist)
    res = {}
    # select -> fetch -> aggregates
    s = t.select("c1", -3000, 4000)
    f = t.fetch("c2", s)
    res["agg"] = t.aggregate(f)
    res["pos"] = t.gather_global(s.local, s.base).numpy()
    res["off"] = (s.offset, s.total, s.local.numel())
    res["empty"] = t.aggregate(t.fetch("c2", t.select("c1", 5, 5)))
    # add / sub stay local
    g = t.fetch("c1", s)
    res["add"] = t.gather_global(t.add(f, g)).numpy()
    # batched shared scan: one count exchange for the whole batch
    lows, highs = [-100, 0, 9000, 7], [100, 2500, 9100, 3]
    ss = t.shared_select("c1", lows, highs)
    res["ss"] = [t.gather_global(p.local, p.base).numpy() for p in ss]
    res["ss_off"] = [(p.offset, p.total) for p in ss]
    # per-shard index: shard-major, value order inside a shard
    t.build_index("k")
    si = t.select_index("k", 100, 140)
    res["ix"] = t.gather_global(si.local, si.base).numpy()
    res["ix_fetch"] = t.gather_global(t.fetch("k", si)).numpy()
    # hash join of two prefiltered sides: global positions travel with the keys
    s1, s2 = t.select("c1", None, 9000), t.select("c1", -2000, -1500)
    v1, v2 = t.fetch("k", s1), t.fetch("k", s2)
    p1, p2 = s1.local + s1.base, s2.local + s2.base
    o1, o2 = t.hash_join(v1, p1, v2, p2)
    res["join"] = np.stack([t.gather_global(o1).numpy(), t.gather_global(o2).numpy()], 1)
    res["join_local"] = o1.numel()
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_shards(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gloo") / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    return torch.load(out, weights_only=False)


def test_select_fetch_aggregate_equals_one_shard(two_shards, port):
    tab = _table()
    pos = port.select_scan(tab["c1"], -3000, 4000)
    vals = port.fetch(tab["c2"], pos)
    a = two_shards["agg"]
    assert np.array_equal(two_shards["pos"], pos.astype(np.int64))
    assert (a["sum"], a["count"], a["min"], a["max"]) == (port.sum(vals), pos.size, port.min(vals), port.max(vals))
    assert a["avg"] == port.avg(vals)                      # same two casts + one fp64 divide
    off, total, local = two_shards["off"]
    assert off == 0 and total == pos.size and 0 < local < total
    e = two_shards["empty"]
    assert (e["sum"], e["count"], e["min"], e["max"]) == (0, 0, 2**31 - 1, -2**31) and np.isnan(e["avg"])
    assert np.array_equal(two_shards["add"], port.add(vals, port.fetch(tab["c1"], pos)).astype(np.int64))


def test_shared_select_equals_one_shard(two_shards, port):
    tab = _table()
    exp = port.shared_select(tab["c1"], [-100, 0, 9000, 7], [100, 2500, 9100, 3])
    for got, e, (off, total) in zip(two_shards["ss"], exp, two_shards["ss_off"]):
        assert np.array_equal(got, e.astype(np.int64)) and off == 0 and total == e.size


def test_sharded_index_select_equals_scan_as_a_set(two_shards, port):
    tab = _table()
    exp = port.select_scan(tab["k"], 100, 140)
    assert np.array_equal(np.sort(two_shards["ix"]), exp.astype(np.int64))
    v = two_shards["ix_fetch"]
    half = int(np.searchsorted(two_shards["ix"] >= (N // 2), True))      # shard-major ...
    assert np.all(np.diff(v[:half]) >= 0) and np.all(np.diff(v[half:]) >= 0)   # ... value order inside


def test_partitioned_join_equals_one_shard(two_shards, port):
    tab = _table()
    s1, s2 = port.select_scan(tab["c1"], None, 9000), port.select_scan(tab["c1"], -2000, -1500)
    e1, e2 = port.hash_join(port.fetch(tab["k"], s1), s1, port.fetch(tab["k"], s2), s2)
    exp = np.stack([e1, e2], 1).astype(np.int64)
    got = two_shards["join"]
    assert got.shape == exp.shape and 0 < two_shards["join_local"] < exp.shape[0]

    def canon(a):
        return a[np.lexsort((a[:, 0], a[:, 1]))]
    assert np.array_equal(canon(got), canon(exp))


def test_shard_ranges_tile_the_table():
    from analytical_database_b200.sharded import shard_range
    for n in (0, 1, 7, 4_000_000_000):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, g, w) for g in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_route_hash_twin():
    """The numpy twin of adb_route_pairs' destination function (checked against the CUDA
    kernel in tests/test_gpu_sharded.py) spreads keys and keeps equal keys together."""
    from sharded_cpu_ops import route_dest
    keys = np.arange(1, 100001, dtype=np.int32)
    for parts in (1, 2, 4, 8):
        d = route_dest(keys, parts)
        c = np.bincount(d, minlength=parts)
        assert d.min() >= 0 and d.max() < parts and c.min() > 0.8 * keys.size / parts
    assert np.array_equal(route_dest(np.array([-5, -5, 7], np.int32), 8)[:2], route_dest(np.array([-5, -5], np.int32), 8))
