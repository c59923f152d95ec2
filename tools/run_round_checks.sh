# usage: bash tools/run_round_checks.sh TAG   (inside gpurun, one GPU)
# What the driver runs at round end (GPU tests, smoke, default bench, reference arm), then the
# ncu evidence: the launch list of the same bench command and full captures of the chain
# kernels (materialised and unmaterialised forms).
T=${1:-rXX}
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -c 600 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -2 gpurun_out/${T}_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["chain"]["frac_of_aggregate_peak"], d["roofline"]["frac"], d["e2e"]["value"], d.get("e2e_cold",{}).get("value"), d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print(json.dumps(d.get("chain_unmaterialised"))[:600])
print(json.dumps(d.get("hash_join"))[:1200])
print(json.dumps(d.get("config2_shared_scan"))[:800])
print(json.dumps(d.get("config3_index"))[:1600])
PY
# full captures: (1) the materialised chain's two kernels, (2) the unmaterialised forms
CH="python bench.py --steps 2 --warmup 3 --shards-limit 1 --no-cpu --no-cold --no-sweep --no-join --no-configs --no-lazy"
timeout 600 $CH > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mask_kernel|expand_kernel" -s 4 -c 2 -o gpurun_out/${T}_chain -f $CH > gpurun_out/${T}_ncu1.log 2>&1; tail -1 gpurun_out/${T}_ncu1.log
timeout 300 python tools/lazy_probe.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_gather_agg_kernel|bitmap_gather_agg_kernel" -s 60 -c 2 -o gpurun_out/${T}_lazy -f python tools/lazy_probe.py > gpurun_out/${T}_ncu3.log 2>&1; tail -1 gpurun_out/${T}_ncu3.log
# every launch of the default bench command's device-resident part, with its device time
BL="python bench.py --steps 2 --warmup 3 --no-cpu --no-cold --no-sweep --no-join --no-configs"
timeout 600 $BL > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $BL > gpurun_out/${T}_ncu2.log 2>&1; tail -1 gpurun_out/${T}_ncu2.log
