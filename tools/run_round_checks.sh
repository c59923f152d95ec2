# usage: bash tools/run_round_checks.sh TAG   (inside gpurun, one GPU)
# What the driver runs at round end (GPU tests, smoke, default bench, reference arm), then the
# ncu evidence of the same bench command: launch list + one full capture of the chain kernels.
T=${1:-rXX}
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -c 600 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --ops > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -2 gpurun_out/${T}_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["chain"]["frac_of_aggregate_peak"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("cold",{}).get("value"), d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print(json.dumps(d.get("hash_join"))[:800])
print(json.dumps(d.get("ops",{}))[:3000])
PY
timeout 600 python bench.py --steps 2 --warmup 3 --shards-limit 1 --no-cpu --no-cold --no-sweep --no-join > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mask_kernel|expand_kernel" -s 4 -c 2 -o gpurun_out/${T}_chain -f python bench.py --steps 2 --warmup 3 --shards-limit 1 --no-cpu --no-cold --no-sweep --no-join > gpurun_out/${T}_ncu1.log 2>&1; tail -1 gpurun_out/${T}_ncu1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-cold --no-sweep --no-join > gpurun_out/${T}_ncu2.log 2>&1; tail -1 gpurun_out/${T}_ncu2.log
