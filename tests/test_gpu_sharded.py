"""GPU tests of the sharded path: the routing kernel against its numpy twin, the sharded
operators at world size 1 against the oracle, and -- when the box has >= 2 GPUs -- a real
NCCL run (torchrun, one process per GPU) whose concatenated results must equal the oracle's
single-shard results."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def eng():
    from analytical_database_b200 import Engine
    e = Engine(0)
    yield e


@pytest.mark.parametrize("parts", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [0, 1, 1000, 300_007])
def test_route_pairs_matches_the_twin(eng, rng, parts, n):
    import ctypes as C
    from sharded_cpu_ops import route_dest
    val = rng.integers(-50_000, 50_000, n).astype(np.int32)
    pos = rng.permutation(n).astype(np.int32)
    dv, dp = eng.upload(val), eng.upload(pos)
    ov, op = eng.alloc_i32(n), eng.alloc_i32(n)
    counts = (C.c_int64 * parts)()
    eng._ck(eng.lib.adb_route_pairs(dv.i32(), dp.i32(), n, parts, ov.i32(), op.i32(), counts))
    d = route_dest(val, parts)
    order = np.argsort(d, kind="stable")
    assert list(counts) == np.bincount(d, minlength=parts).tolist()
    assert np.array_equal(ov.to_host(n), val[order]) and np.array_equal(op.to_host(n), pos[order])


def test_sharded_world1_equals_oracle(eng, port, rng):
    from analytical_database_b200.sharded import EngineOps, ShardedTable
    from sharded_gpu_worker import table
    n = 200_003
    tab = table(n)
    dev = torch.device("cuda", 0)
    t = ShardedTable(EngineOps(eng, dev), {k: torch.from_numpy(v).to(dev) for k, v in tab.items()}, n)
    s = t.select("c1", -n // 20, n // 10)
    f = t.fetch("c2", s)
    a = t.aggregate(f)
    pos = port.select_scan(tab["c1"], -n // 20, n // 10)
    vals = port.fetch(tab["c2"], pos)
    assert np.array_equal(s.local.cpu().numpy(), pos)
    assert (a["sum"], a["count"], a["min"], a["max"], a["avg"]) == \
        (port.sum(vals), pos.size, port.min(vals), port.max(vals), port.avg(vals))
    o1, o2 = t.hash_join(t.fetch("k", s), s.local, t.fetch("k", s), s.local)
    e1, e2 = port.hash_join(port.fetch(tab["k"], pos), pos, port.fetch(tab["k"], pos), pos)
    assert np.array_equal(o1.cpu().numpy(), e1) and np.array_equal(o2.cpu().numpy(), e2)
    eng.set_stream(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_nccl_shards_equal_one_shard(port, tmp_path):
    from sharded_gpu_worker import table
    world = 2 if torch.cuda.device_count() < 4 else 4
    n = 2_000_003
    out = str(tmp_path / "res.pt")
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        p = s_.getsockname()[1]
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                    f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(p),
                    os.path.join(HERE, "sharded_gpu_worker.py"), out, str(n)], check=True, timeout=600)
    res = torch.load(out, weights_only=False)
    tab = table(n)
    pos = port.select_scan(tab["c1"], -n // 20, n // 10)
    vals = port.fetch(tab["c2"], pos)
    a = res["agg"]
    assert res["world"] == world and np.array_equal(res["pos"], pos.astype(np.int64))
    assert (a["sum"], a["count"], a["min"], a["max"], a["avg"]) == \
        (port.sum(vals), pos.size, port.min(vals), port.max(vals), port.avg(vals))
    assert all(x == a for x in res["agg_peer"])           # peer-memory exchange == NCCL exchange
    for x in res["agg_fused"]:                              # chain kernel with the exchange as its epilogue
        assert (x["sum"], x["count"], x["min"], x["max"]) == (a["sum"], a["count"], a["min"], a["max"])
    e = res["agg_peer_empty"]
    assert (e["sum"], e["count"]) == (0, 0) and e["avg"] != e["avg"]
    exp = port.shared_select(tab["c1"], [-100, 0, n // 4, 7], [100, n // 16, n // 4 + n // 50, 3])
    for got, e in zip(res["ss"], exp):
        assert np.array_equal(got, e.astype(np.int64))
    scan = port.select_scan(tab["k"], 100, 140).astype(np.int64)
    assert np.array_equal(np.sort(res["ix"]), scan) and np.array_equal(res["ix"], res["ix_tree"])
    s1, s2 = port.select_scan(tab["c1"], None, n // 4), port.select_scan(tab["c1"], -n // 10, -n // 20)
    e1, e2 = port.hash_join(port.fetch(tab["k"], s1), s1, port.fetch(tab["k"], s2), s2)
    expj = np.stack([e1, e2], 1).astype(np.int64)

    def canon(x):
        return x[np.lexsort((x[:, 0], x[:, 1]))]
    assert res["join"].shape == expj.shape and np.array_equal(canon(res["join"]), canon(expj))
    assert 0 < res["join_local"] < expj.shape[0] and res["launches"] > 20
    for got in res["join_peer"]:                           # pair exchange over NVLink peer memory
        assert got.shape == expj.shape and np.array_equal(canon(got), canon(expj))
    assert res["xchg_equal"]                               # same pieces, same order as the all-to-all-v
    assert "reserve more" in res["overflow"] and res["after_overflow"] == 8 * world
