"""torchrun worker for tests/test_gpu_sharded.py: the sharded operator path on real GPUs over
NCCL, one process per GPU.  Rank 0 saves what the test compares with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def table(n, seed=42):
    rng = np.random.default_rng(seed)
    return {"c1": rng.integers(-n // 2, n // 2, n).astype(np.int32),
            "c2": rng.integers(2**31 - 10000, 2**31 - 1, n, dtype=np.int64).astype(np.int32),
            "k": rng.integers(1, n // 8, n).astype(np.int32)}


def main():
    out, n = sys.argv[1], int(sys.argv[2])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import analytical_database_b200 as adb
    from analytical_database_b200.sharded import EngineOps, ShardedTable, shard_range
    eng = adb.Engine(local)
    ops = EngineOps(eng, dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    tab = table(n)
    b, e = shard_range(n, rank, world)
    t = ShardedTable(ops, {k: torch.from_numpy(v[b:e].copy()).to(dev) for k, v in tab.items()}, n, dist)
    res = {"world": world}
    s = t.select("c1", -n // 20, n // 10)
    f = t.fetch("c2", s)
    res["agg"] = t.aggregate(f)                       # two NCCL all-reduces
    ops.connect_peers(dist)
    res["agg_peer"] = [t.aggregate(f) for _ in range(5)]      # one kernel over NVLink peer memory
    res["agg_peer_empty"] = t.aggregate(f[:0])
    # the chain with the exchange fused into its last kernel: this rank's rows as two shards
    import ctypes as C
    from analytical_database_b200.engine import _AggStruct
    rows = e - b
    half = rows // 2
    c1, c2 = t.cols["c1"], t.cols["c2"]
    pos, val = ops.empty(rows), ops.empty(rows)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    parts = torch.zeros(3 * 2, dtype=torch.int64, device=dev)          # 2 x adb_agg (24 bytes each)
    agg_out = torch.zeros(3, dtype=torch.int64, device=dev)
    P = lambda x, ty=C.c_int32, off=0: C.cast(C.c_void_p(x.data_ptr() + off), C.POINTER(ty))
    lo, hi = C.c_int32(-n // 20), C.c_int32(n // 10)
    res["agg_fused"] = []
    for _ in range(3):
        eng._ck(eng.lib.adb_chain_select_fetch_agg(P(c1), P(c2), half, C.byref(lo), C.byref(hi), P(pos), P(val),
                                                   P(cnt, C.c_int64), P(parts, _AggStruct)))
        eng._ck(eng.lib.adb_chain_select_fetch_agg_exchange(
            P(c1, off=4 * half), P(c2, off=4 * half), rows - half, C.byref(lo), C.byref(hi), P(pos), P(val),
            P(cnt, C.c_int64), P(parts, _AggStruct), 2, P(agg_out, _AggStruct)))
        torch.cuda.synchronize()
        h = _AggStruct.from_buffer_copy(agg_out.cpu().numpy().tobytes())
        res["agg_fused"].append({"sum": h.sum, "count": h.count, "min": h.min, "max": h.max})
    res["pos"] = t.gather_global(s.local, s.base).cpu().numpy()
    res["off"] = (s.offset, s.total, s.local.numel())
    lows = [-100, 0, n // 4, 7]
    highs = [100, n // 16, n // 4 + n // 50, 3]
    ss = t.shared_select("c1", lows, highs)
    res["ss"] = [t.gather_global(p.local, p.base).cpu().numpy() for p in ss]
    t.build_index("k")
    for tree in (False, True):
        si = t.select_index("k", 100, 140, use_btree=tree)
        res["ix_tree" if tree else "ix"] = t.gather_global(si.local, si.base).cpu().numpy()
    s1, s2 = t.select("c1", None, n // 4), t.select("c1", -n // 10, -n // 20)
    v1, v2 = t.fetch("k", s1), t.fetch("k", s2)
    p1, p2 = s1.local + s1.base, s2.local + s2.base
    o1, o2 = t.hash_join(v1, p1, v2, p2)
    res["join"] = np.stack([t.gather_global(o1).cpu().numpy(), t.gather_global(o2).cpu().numpy()], 1)
    res["join_local"] = o1.numel()
    # the same join with the pair exchange over peer memory (three times: banks, buffer reuse)
    ops.connect_peers(dist, join_cap_pairs=n)
    res["join_peer"] = []
    for _ in range(3):
        q1, q2 = t.hash_join(v1, p1, v2, p2)
        res["join_peer"].append(np.stack([t.gather_global(q1).cpu().numpy(), t.gather_global(q2).cpu().numpy()], 1))
    # and the exchange alone against the NCCL all-to-all-v: identical pieces in identical order
    ops.peer_join = False
    nv, npos = t.exchange_pairs(v2, p2)
    ops.peer_join = True
    pv, ppos = t.exchange_pairs(v2, p2, 1)
    ok = torch.tensor([int(torch.equal(nv, pv) and torch.equal(npos, ppos))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)                 # on every rank
    res["xchg_equal"] = bool(ok.item())
    # a receive region that is too small fails on every rank alike, and the next exchange works
    ops.connect_peers(dist, join_cap_pairs=64)
    try:
        t.exchange_pairs(v1, p1, 0)
        res["overflow"] = "no error"
    except Exception as ex:                                   # EngineError: ADB_ERR_NOMEM
        res["overflow"] = str(ex)
    small = t.exchange_pairs(v1[:8], p1[:8], 0)
    res["after_overflow"] = int(t.gather_global(small[0]).numel())
    res["launches"] = eng.launch_count()
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
