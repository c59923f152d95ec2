#!/usr/bin/env python
"""bench.py -- the north-star measurement: rows/s for select -> fetch -> sum/min/max over
the 4 B-row int32 table of BASELINE.json config 5, row-range partitioned into 8 shards of
500 M rows (positions are int32, /root/reference/src/query.c:94-95, so a shard must stay
below 2^31 rows) spread over N B200s, aggregate partials combined with an NCCL allreduce.

One "step" = one pass of the chain over every shard of the table on every rank:
  s = select(tbl.col1, lo, hi)      adb_select_scan   (query.c:92)
  f = fetch(tbl.col2, s)            adb_fetch         (query.c:223)
  a = sum(f) / min(f) / max(f)      adb_aggregate     (query.c:325,392,417)
with the position list and the fetched vector materialised, as the operator API defines.

  python bench.py [--gpus N] [--steps K] [--warmup W]            engine arm
  python bench.py --impl reference ...                           reference CPU arm
N > 1 is launched by torchrun (one rank per GPU); total work is fixed (strong scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_ROWS = 4_000_000_000
N_SHARDS = 8
SEED = 42
SPAN = 1 << 30                       # col1 uniform in [0, 2^30)
FETCH_LO, FETCH_SPAN = 2**31 - 10000, 10000     # col2 near INT_MAX (milestone1.py:119)
METRIC = "rows/sec for select+fetch+sum"
UNIT = "rows/s"


def predicate(selectivity: float):
    lo = 1000
    return lo, lo + int(SPAN * selectivity)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region, in-process through NVML
    (a sample every ~2 ms; the timed region is tens of milliseconds, too short for an
    `nvidia-smi -lms` child to report from).  Falls back to one nvidia-smi query."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
            "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            self.reasons |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            try:
                self.reasons |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass

    def _loop(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.002)

    def start(self):
        if self.nvml is None:
            return
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self._thread.join()
            if not self.sm:
                self._sample()
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, b in self.BITS.items() if self.reasons & b),
                    "samples": len(self.sm), "source": "nvml, sampled inside the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 stdout=subprocess.PIPE, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [],
                    "samples": 1, "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU operators on the host cores
# ------------------------------------------------------------------------------------------
def host_columns(rows: int, first_row: int, threads: int):
    """The same rows the GPU scans, regenerated on the host (counter-based generator)."""
    from analytical_database_b200 import synth
    from concurrent.futures import ThreadPoolExecutor
    sel = np.empty(rows, np.int32)
    fet = np.empty(rows, np.int32)
    chunk = 1 << 24

    def fill(b):
        e = min(rows, b + chunk)
        sel[b:e] = synth.uniform(e - b, SEED, first_row + b, 0, SPAN)
        fet[b:e] = synth.uniform(e - b, SEED + 1, first_row + b, FETCH_LO, FETCH_SPAN)

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(fill, range(0, rows, chunk)))
    return sel, fet


def cpu_ops():
    from oracle import oracle
    ref = oracle.reference("O2")
    if ref is not None:
        return ref, "reference", "oracle/_ref/libref_O2.so = unmodified reference query.c at -O2"
    return oracle.port(), "port", "oracle/liboracle.so restatement at -O2"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    ops, kind, what = cpu_ops()
    lo, hi = predicate(args.selectivity)
    rows = args.cpu_rows
    sel, fet = host_columns(rows, 0, cores)
    for _ in range(args.warmup):
        ops.chain_select_fetch_sum(sel, fet, lo, hi, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, h = ops.chain_select_fetch_sum(sel, fet, lo, hi, threads=cores)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32 (int64 accumulate)", "data": "synthetic",
        "config": workload_config(args, rows_per_step=rows),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{rows} rows of the same table per step, one reference "
                                   f"instance per row range on {cores} threads; {what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "check": {"sum": s, "hits": h},
    }
    emit(line)
    return 0


def workload_config(args, rows_per_step):
    return {
        "workload": "BASELINE config 5: 4B-row multi-column int32 table row-partitioned "
                    "across 1/2/4/8 B200: select+fetch+sum/min/max with NCCL allreduce",
        "total_rows": TOTAL_ROWS, "shards": N_SHARDS, "rows_per_shard": TOTAL_ROWS // N_SHARDS,
        "rows_per_step": rows_per_step, "selectivity": args.selectivity,
        "columns": "col1 uniform [0,2^30) (select), col2 uniform [2^31-10000,2^31) (fetch)",
        "l2": "inputs exceed L2 (>= 2 GB per shard column vs 126 MB)",
    }


# ------------------------------------------------------------------------------------------
# engine arm
# ------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import analytical_database_b200 as adb
    from analytical_database_b200.engine import AGG_BYTES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = adb.Engine(local)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)          # engine kernels and NCCL share one stream
    torch.cuda.set_stream(stream)

    if dist is not None and args.exchange == "peer":
        eng.peer_setup(dist)                    # map every rank's aggregate mailbox (NVLink P2P)
    if dist is None and args.exchange_self:     # debug: the exchange kernel variant at world size 1
        hbuf = C.create_string_buffer(64)
        eng._ck(eng.lib.adb_peer_create(1, 0, hbuf))
        eng._ck(eng.lib.adb_peer_connect(hbuf.raw))

    lo, hi = predicate(args.selectivity)
    shard_rows = TOTAL_ROWS // N_SHARDS
    my_shards = [s for s in range(N_SHARDS) if s % world == rank] if world > 1 else list(range(N_SHARDS))
    if args.shards_limit:
        my_shards = my_shards[:args.shards_limit]
    # ---- load: columns become HBM-resident int32 arrays -------------------------------
    cols = []
    for s in my_shards:
        c1 = eng.synth_uniform(shard_rows, SEED, s * shard_rows, 0, SPAN)
        c2 = eng.synth_uniform(shard_rows, SEED + 1, s * shard_rows, FETCH_LO, FETCH_SPAN)
        cols.append((c1, c2))
    cap = int(shard_rows * min(1.0, args.selectivity * 1.5 + 0.001)) + 4096
    # one (pos, val, count) result set per shard: every handle stays materialised
    res = [(eng.alloc_i32(cap), eng.alloc_i32(cap), eng.alloc(8)) for _ in my_shards]
    parts = eng.alloc(AGG_BYTES * max(len(my_shards), 1))
    # the allreduce operands live in torch tensors (plumbing for NCCL)
    t_sum = torch.zeros(2, dtype=torch.int64, device="cuda")      # {sum, count}
    t_mm = torch.zeros(2, dtype=torch.int32, device="cuda")       # {max, ~min}
    combined = eng.alloc(AGG_BYTES)
    blo, bhi = C.c_int32(lo), C.c_int32(hi)
    lib = eng.lib
    AggP = eng.agg_ptr

    def step(mark_base=None):
        fused_exchange = (dist is not None and args.exchange == "peer") or (dist is None and args.exchange_self)
        for i, (c1, c2) in enumerate(cols):
            pos, val, cnt = res[i]
            if mark_base is not None:
                eng._ck(lib.adb_chain_marks(mark_base + 3 * i))
            # predicate pass + expansion with the gather and the aggregates fused in; the
            # position list and the fetched vector are still materialised (pos, val)
            if fused_exchange and i == len(cols) - 1:
                # the rank's last shard: the CTA that completes its aggregate also folds the
                # earlier shards' partials and swaps the result with every peer over NVLink
                # peer memory, inside the same kernel (csrc/peer_agg.cu, adb_common.cuh)
                eng._ck(lib.adb_chain_select_fetch_agg_exchange(
                    c1.i32(), c2.i32(), shard_rows, C.byref(blo), C.byref(bhi), pos.i32(), val.i32(),
                    cnt.i64(), AggP(parts), len(cols), AggP(combined)))
            else:
                eng._ck(lib.adb_chain_select_fetch_agg(c1.i32(), c2.i32(), shard_rows, C.byref(blo),
                                                       C.byref(bhi), pos.i32(), val.i32(), cnt.i64(),
                                                       AggP(parts, i)))
        if fused_exchange:
            pass
        elif dist is None:
            eng._ck(lib.adb_agg_combine(AggP(parts), len(cols), AggP(combined), None))
        else:
            eng._ck(lib.adb_agg_combine(AggP(parts), len(cols), AggP(combined), None))
            eng._ck(lib.adb_agg_export(AggP(combined), C.c_void_p(t_sum.data_ptr()),
                                       C.c_void_p(t_mm.data_ptr())))
            dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
            dist.all_reduce(t_mm, op=dist.ReduceOp.MAX)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # CUDA events between the chain's two launches give the live per-kernel times for the
    # roofline object, but an event record between two kernels also keeps the second one from
    # being scheduled while the first drains (programmatic dependent launch), so only every
    # fourth timed step carries them
    marks_per_step = 3 * len(cols)
    MARK_EVERY = 4
    marked_steps = [k for k in range(args.steps) if k % MARK_EVERY == 0]
    timed_marks = len(marked_steps) * marks_per_step <= 8000
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step((k // MARK_EVERY) * marks_per_step if timed_marks and k % MARK_EVERY == 0 else None)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    rows_step = shard_rows * (len(my_shards) * world if args.shards_limit else N_SHARDS)
    ms_step = ms_total / args.steps
    value = rows_step / (ms_step * 1e-3)

    # ---- result of the last step (device-resident) -----------------------------------------
    if dist is not None and args.exchange == "nccl":
        g_sum, g_cnt = int(t_sum[0].item()), int(t_sum[1].item())
        g_max, g_min = int(t_mm[0].item()), ~int(t_mm[1].item())
    else:
        a = eng.read_agg(combined)
        g_sum, g_cnt, g_min, g_max = a.sum, a.count, a.min, a.max

    # ---- roofline of the dominant kernel (mask_kernel: the predicate pass over the selected
    # column), from CUDA events recorded on the engine stream inside the timed region -------
    peak, peak_src = measured_peak()
    hits_local = [int(r[2].to_host(1, np.int64)[0]) for r in res]
    mask_ms, fused_ms = [], []
    if timed_marks:
        for k in range(len(marked_steps)):
            for i in range(len(cols)):
                b = k * marks_per_step + 3 * i
                mask_ms.append(eng.mark_elapsed(b, b + 1))
                fused_ms.append(eng.mark_elapsed(b + 1, b + 2))
    roofline = None
    if mask_ms:
        avg_ms = float(np.mean(mask_ms))
        alg_bytes = 4.0 * shard_rows                     # the 4N of select's 4N + 4H
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "select_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        h_avg = float(np.mean(hits_local))
        f_ms = float(np.mean(fused_ms))
        roofline = {"bound": "hbm", "kernel": "adb::mask_kernel (predicate pass of adb_select_scan / "
                    "adb_chain_select_fetch_agg: column -> 1 bit/row bitmap + per-chunk counts)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": avg_ms,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "share_of_step": float(np.mean(mask_ms) * len(cols) / ms_step),
                    "second_kernel": {
                        "kernel": "adb::expand_kernel<false,true> (bitmap -> positions, gather, "
                                  "sum/min/max fused)",
                        "avg_launch_ms": f_ms, "share_of_step": float(f_ms * len(cols) / ms_step),
                        "algorithmic_bytes_per_launch": 16.0 * h_avg,
                        "achieved": 16.0 * h_avg / (f_ms * 1e-3) / 1e9,
                        "miss_granular_bytes_per_launch": 64.0 * h_avg + 8.0 * h_avg + shard_rows / 8.0,
                        "note": "a sparse gather moves 64 bytes per hit with ld.global.nc.L2::64B "
                                "(128 with a plain load; profiles/r01c_gather_probe.md): the kernel's "
                                "DRAM traffic is about 5x its algorithmic bytes, and at 5 M random "
                                "reads per launch it runs at the DRAM random-access rate"}}
    chain_bytes = 4.0 * rows_step + 20.0 * g_cnt
    chain_gbs = chain_bytes / (ms_step * 1e-3) / 1e9

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, 1),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 (int64 accumulate)", "data": "synthetic",
            "config": workload_config(args, rows_step),
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "chain": {"algorithmic_bytes_per_step": chain_bytes, "achieved_gbs": chain_gbs,
                      "frac_of_aggregate_peak": chain_gbs / (peak * max(world, 1)),
                      "formula": "4N + 20H (SURVEY.md 8d)"},
            "result": {"sum": g_sum, "count": g_cnt, "min": g_min, "max": g_max},
        }
        if world > 1:
            line["config"]["aggregate_exchange"] = (
                "adb_chain_select_fetch_agg_exchange: the chain kernel of the rank's last shard folds the partials and exchanges them over NVLink peer memory (no separate launch, no NCCL call)"
                if args.exchange == "peer" else "adb_agg_export + 2 NCCL all-reduces")

    # ---- selectivity sweep on one shard (SURVEY.md 8d lists 0.1 %, 1 %, 10 %, 50 %) -----------
    if rank == 0 and not args.no_sweep:
        sweep = {}
        c1, c2 = cols[0]
        for sel in (0.001, 0.01, 0.1, 0.5):
            slo, shi = predicate(sel)
            need = int(shard_rows * min(1.0, sel * 1.05 + 0.001)) + 4096
            sp, sv = eng.alloc_i32(need), eng.alloc_i32(need)
            b1, b2 = C.c_int32(slo), C.c_int32(shi)

            def one():
                eng._ck(lib.adb_chain_select_fetch_agg(c1.i32(), c2.i32(), shard_rows, C.byref(b1),
                                                       C.byref(b2), sp.i32(), sv.i32(), res[0][2].i64(),
                                                       AggP(parts, 0)))
            for _ in range(2):
                one()
            eng.timer_start()
            for _ in range(5):
                one()
            ms = eng.timer_stop() / 5
            h = int(res[0][2].to_host(1, np.int64)[0])
            gbs = (4.0 * shard_rows + 20.0 * h) / (ms * 1e-3) / 1e9
            sweep[str(sel)] = {"ms_per_shard": ms, "hits": h, "rows_per_s_per_gpu": shard_rows / (ms * 1e-3),
                               "chain_algorithmic_gbs": gbs, "frac_of_peak": gbs / peak}
            sp.free()
            sv.free()
        line["selectivity_sweep"] = sweep

    # second half of BASELINE.json's metric: hash-join tuples/s (config 4).  N > 1: both tables
    # row-range sharded, pairs hash-routed to their owners, joined locally.
    join_sharded = None
    if world > 1 and not args.no_join:
        join_sharded = measure_join_sharded(eng, dist, rank, world, local)

    # ---- e2e and cpu_baseline (rank 0; the CPU leg only at N = 1) ----------------------------
    e2e = measure_e2e(args, eng, cols, res, shard_rows, lo, hi, world, dist, rank, t_sum, t_mm,
                      combined, parts)
    if rank == 0:
        line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = measure_cpu(args, eng, cols, res, shard_rows, lo, hi)
        if world == 1 and (args.ops or not args.no_join):
            for c1, c2 in cols[1:]:
                c1.free()
                c2.free()
        if world == 1 and not args.no_join and not args.ops:
            line["hash_join"] = measure_join_single(args)
        if world == 1 and args.ops:
            line["ops"] = measure_ops(args)
            if "hash_join_100Mx100M" in line["ops"]:
                line["hash_join"] = join_summary(line["ops"]["hash_join_100Mx100M"], 1)
        if world > 1 and join_sharded is not None:
            line["hash_join"] = join_summary(join_sharded, world)
        emit(line)
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def measure_e2e(args, eng, cols, res, shard_rows, lo, hi, world, dist, rank, t_sum, t_mm,
                combined, parts):
    """The same chain through the reference-facing operator API -- the C host drop-in
    (host/query_shim.c -> libadb_query.so) called exactly as src/server.c:137-290 calls
    query.c: select_column(Column*, &lo, &hi) -> fetch_column(Column*, Result*) ->
    sum(GeneralizedColumn*) with host Column / Result / Status structs, every operator
    returning its num_tuples to the host before the next is issued (parse.c:799 needs it),
    the long read back from the scalar Result, and the handles released the way
    client_context.c does.  Wall clock.

    warm: a table is uploaded to HBM once, when the shim first touches it (the reference
          keeps loaded columns in RAM the same way); the timed steps then move only the
          bounds down and the counts / sum up.
    cold: one shard whose two columns live in HOST memory and are invalidated before every
          step, so each step pays the H2D copy of 2 x 2 GB inside the timed region."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import query_api as q                      # ctypes view of include/adb_query_api.h
    api = q.Api()
    L = api.lib

    def make_col(devbuf, host=None):
        c = q.Column()
        c.name = b"col"
        c.row_count = shard_rows
        if host is not None:
            c.data = host.ctypes.data_as(C.POINTER(C.c_int))
        return c

    hcols = []
    for (c1, c2) in cols:
        a, b = make_col(c1), make_col(c2)
        assert L.adb_host_column_adopt(C.byref(a), c1.void()) == 0
        assert L.adb_host_column_adopt(C.byref(b), c2.void()) == 0
        hcols.append((a, b))
    blo, bhi = C.c_int(lo), C.c_int(hi)
    st = q.Status(99, None)
    gen = q.GeneralizedColumn(q.RESULT)

    def chain(a, b):
        s_ = L.select_column(C.byref(a), C.byref(blo), C.byref(bhi), C.byref(st))
        if st.code != q.OK:
            raise SystemExit("select_column: " + L.adb_host_last_error().decode())
        f_ = L.fetch_column(C.byref(b), s_, C.byref(st))
        gen.column_pointer.result = f_
        r_ = L.sum(C.byref(gen), C.byref(st))
        if st.code != q.OK:
            raise SystemExit("fetch/sum: " + L.adb_host_last_error().decode())
        total = C.cast(r_.contents.payload, C.POINTER(C.c_long))[0]
        hits = s_.contents.num_tuples
        for h_ in (s_, f_, r_):
            api.drop(h_)
        return total, hits

    h_pair = torch.zeros(2, dtype=torch.int64).pin_memory()

    def step():
        tot = hits = 0
        for (a, b) in hcols:
            t_, h_ = chain(a, b)
            tot += t_
            hits += h_
        if dist is not None:
            # the ranks' host-side sums meet in one all-reduce: one pinned H2D copy down, one
            # D2H copy back (element-wise tensor writes and .item() cost a launch + sync each)
            h_pair[0], h_pair[1] = tot, hits
            t_sum.copy_(h_pair, non_blocking=True)
            dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
            h_pair.copy_(t_sum)
            tot, hits = int(h_pair[0]), int(h_pair[1])
        return tot, hits

    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        tot, hits = step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    rows_step = shard_rows * len(cols) * max(world, 1)
    out = {"value": rows_step * steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": 16 if dist is not None else 0,
           "d2h_bytes_per_step": (8 + 24) * len(cols) + (16 if dist is not None else 0),
           "ms_per_step": 1e3 * dt / steps, "steps": steps,
           "api": "select_column -> fetch_column -> sum of include/adb_query_api.h "
                  "(libadb_query.so = host/query_shim.c), one call chain per shard",
           "mode": "warm: columns HBM-resident after load; bounds travel as kernel arguments "
                   "(no H2D copy); per shard the host reads num_tuples (8 B) after select and "
                   "the aggregate (24 B) after sum; wall clock; see `cold` for the H2D-inclusive figure",
           "check": {"sum": tot, "hits": hits}}
    # cold: HOST columns, re-uploaded by the shim inside every timed step
    if rank == 0 and world == 1 and not args.no_cold:
        c1, c2 = cols[0]
        h1, h2 = c1.to_host(shard_rows), c2.to_host(shard_rows)
        a, b = make_col(c1, h1), make_col(c2, h2)
        chain(a, b)                                         # first touch
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            L.adb_host_column_invalidate(C.byref(a))
            L.adb_host_column_invalidate(C.byref(b))
            ctot, chits = chain(a, b)
        dtc = time.perf_counter() - t0
        L.adb_host_column_invalidate(C.byref(a))
        L.adb_host_column_invalidate(C.byref(b))
        out["cold"] = {"value": shard_rows * reps / dtc, "unit": UNIT,
                       "h2d_bytes_per_step": 8 * shard_rows, "d2h_bytes_per_step": 32,
                       "ms_per_step": 1e3 * dtc / reps,
                       "sample": f"one {shard_rows}-row shard whose two columns are host arrays; the "
                                 "shim uploads both inside every timed step (pageable memory staged "
                                 "through the engine's pinned multi-lane pipeline, adb_upload)",
                       "check": {"sum": ctot, "hits": chits}}
    for (a, b) in hcols:
        L.adb_host_column_invalidate(C.byref(a))
        L.adb_host_column_invalidate(C.byref(b))
    return out


def measure_join_sharded(eng, dist, rank, world, local):
    """BASELINE config 4: hash join of two 100 M-row tables with selective prefilters, both
    row-range sharded over the ranks, pairs hash-routed by key and pushed into the destination
    ranks' memory over NVLink (adb_peer_exchange_pairs), joined locally; the NCCL all-to-all-v
    exchange is timed next to it.  Returns tuples/s over all ranks (max time)."""
    import torch
    from analytical_database_b200.sharded import EngineOps, ShardedTable, shard_range
    dev = torch.device("cuda", local)
    ops = EngineOps(eng, dev)                 # engine stream = torch's current stream
    n = 100_000_000
    out = {}
    b, e = shard_range(n, rank, world)
    rows = e - b

    def col(seed, lo, span):
        t = torch.empty(rows, dtype=torch.int32, device=dev)
        eng._ck(eng.lib.adb_synth_uniform(C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_int32)),
                                          rows, seed, b, lo, span))
        return t
    t1 = ShardedTable(ops, {"k": col(11, 1, n), "f": col(13, 0, 1000)}, n, dist)
    t2 = ShardedTable(ops, {"k": col(12, 1, n), "f": col(14, 0, 1000)}, n, dist)
    # receive regions of the peer-memory pair exchange: 30 % head room over an even split
    ops.connect_peers(dist, join_cap_pairs=int(1.3 * n / world) + (1 << 16))

    def timed_join(v1, g1, v2, g2):
        ms = []
        for it in range(4):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o1, o2 = t1.hash_join(v1, g1, v2, g2)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ms[1:])], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m = torch.tensor([o1.numel()], dtype=torch.int64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.SUM)
        return float(t.item()), int(m.item())

    for s1, s2 in ((0.8, 0.15), (1.0, 1.0)):
        p1, p2 = t1.select("f", None, int(1000 * s1)), t2.select("f", None, int(1000 * s2))
        v1, v2 = t1.fetch("k", p1), t2.fetch("k", p2)
        g1, g2 = p1.local + p1.base, p2.local + p2.base        # global positions (< 2^31)
        tup = p1.total + p2.total
        ops.peer_join = True
        ms_peer, m_peer = timed_join(v1, g1, v2, g2)
        ops.peer_join = False
        ms_nccl, m_nccl = timed_join(v1, g1, v2, g2)
        ops.peer_join = True
        out[f"prefilter_{s1}_{s2}"] = {
            "build": p1.total, "probe": p2.total, "matches": m_peer, "matches_equal": m_peer == m_nccl,
            "ms": ms_peer, "tuples_per_s": tup / (ms_peer * 1e-3), "world": world,
            "includes": "adb_peer_exchange_pairs for both sides (counts all-gather, routing scatter into the "
                        "destination ranks' memory over NVLink, done flags) + local join",
            "nccl_exchange": {"ms": ms_nccl, "tuples_per_s": tup / (ms_nccl * 1e-3),
                              "includes": "adb_route_pairs, 2 x (count + 2 payload) NCCL all-to-all, local join"}}
    return out


def join_summary(cases: dict, world: int) -> dict:
    """The `hash_join` object of the JSON line: tuples/s = (build + probe tuples) / time."""
    out = {"metric": "hash-join tuples/sec", "unit": "tuples/s", "n_gpus": world,
           "workload": "BASELINE config 4: 100M x 100M int32 key columns (uniform in [1, 100M]), prefilter "
                       "select(f, null, x) on each side, hash join of the (value, position) pair lists"
                       + (", pairs hash-partitioned across the ranks over NVLink" if world > 1 else ""),
           "cases": cases}
    full = cases.get("prefilter_1.0_1.0")
    if full:
        out["value"] = full["tuples_per_s"]
        out["ms"] = full["ms"]
    return out


def measure_join_single(args):
    """BASELINE config 4 on one GPU (adb_hash_join_count + adb_join_emit), two prefilter cases."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_ops
    import analytical_database_b200 as adb
    eng = adb.Engine(int(os.environ.get("LOCAL_RANK", "0")))     # same library instance
    try:
        return join_summary(bench_ops.bench_join(eng, 1.0, cases=((0.8, 0.15), (1.0, 1.0))), 1)
    except Exception as e:                                        # the headline line must survive
        return {"error": repr(e)}


def measure_ops(args):
    """The other BASELINE.json configs on one GPU, bounded: batched shared scan (config 2),
    index range select + fetch (config 3), hash join with prefilters (config 4); plus the two
    callers next to the path: bulk CSV load and print of a long result."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_ops
    import analytical_database_b200 as adb
    eng = adb.Engine(int(os.environ.get("LOCAL_RANK", "0")))     # same library instance
    out = {}
    try:
        out["shared_scan_100q_100M"] = bench_ops.bench_shared(eng, 1.0)
        out["index_500M"] = bench_ops.bench_index(eng, 1.0)
        out["hash_join_100Mx100M"] = bench_ops.bench_join(eng, 1.0)
        out["csv_load_4M_rows"] = bench_ops.bench_load(eng, 1.0)
        out["print_50M_values"] = bench_ops.bench_print(eng, 1.0)
    except Exception as e:                                        # the headline line must survive
        out["error"] = repr(e)
    return out


def measure_cpu(args, eng, cols, res, shard_rows, lo, hi):
    """The reference's CPU operators on a bounded sample of the same table, on the box's
    host cores, checked against the GPU's result for the same rows."""
    ops, kind, what = cpu_ops()
    cores = os.cpu_count() or 1
    rows = min(args.cpu_rows, shard_rows)
    c1, c2 = cols[0]
    sel, fet = c1.to_host(rows), c2.to_host(rows)          # the very rows the GPU scanned
    (s1, h1) = ops.chain_select_fetch_sum(sel, fet, lo, hi, threads=1)
    t0 = time.perf_counter()
    ops.chain_select_fetch_sum(sel, fet, lo, hi, threads=1)
    t_single = time.perf_counter() - t0
    reps, t_all = 0, 0.0
    ops.chain_select_fetch_sum(sel, fet, lo, hi, threads=cores)
    t0 = time.perf_counter()
    while t_all < 8.0 and reps < 200:
        s2, h2 = ops.chain_select_fetch_sum(sel, fet, lo, hi, threads=cores)
        reps += 1
        t_all = time.perf_counter() - t0
    # parity: GPU chain on the same rows
    pos, val, cnt = res[0]
    h_cnt = C.c_int64(0)
    blo, bhi = C.c_int32(lo), C.c_int32(hi)
    eng._ck(eng.lib.adb_select_scan(c1.i32(), rows, C.byref(blo), C.byref(bhi), 0, pos.i32(),
                                    cnt.i64(), C.byref(h_cnt)))
    eng.fetch(c2, pos, h_cnt.value, out=val)
    g = eng.aggregate(val, h_cnt.value)
    ok = (g.sum, h_cnt.value) == (s1, h1) == (s2, h2)
    out = {"value": rows * reps / t_all, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"first {rows} rows of shard 0 (downloaded from HBM: identical rows), "
                     f"{reps} passes on {cores} threads (one reference instance per row range); {what}",
           "single_thread_value": rows / t_single, "parity_with_gpu": bool(ok)}
    from oracle import oracle
    r0 = oracle.reference("O0")
    if r0 is not None:
        t0 = time.perf_counter()
        r0.chain_select_fetch_sum(sel, fet, lo, hi, threads=1)
        out["single_thread_O0_value"] = rows / (time.perf_counter() - t0)
    if not ok:
        raise SystemExit(f"PARITY FAILURE: gpu {(g.sum, h_cnt.value)} cpu {(s1, h1)}")
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line
    on stdout.  Point fd 1 at stderr for the run and keep the real stdout for the result."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--selectivity", type=float, default=0.01)
    ap.add_argument("--cpu-rows", type=int, default=500_000_000)
    ap.add_argument("--shards-limit", type=int, default=0, help="debug: fewer shards per rank")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cold", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-join", action="store_true", help="skip the hash-join measurement (config 4)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: aggregate exchange through the engine's peer-memory kernel or NCCL")
    ap.add_argument("--exchange-self", action="store_true",
                    help="debug (N = 1): run the last shard through the exchange-carrying chain kernel")
    ap.add_argument("--ops", action="store_true",
                    help="also time shared scan / index / join at the BASELINE config sizes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_engine(args)


if __name__ == "__main__":
    sys.exit(main())
