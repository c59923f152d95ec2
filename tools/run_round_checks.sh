set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r01m_pytest.log 2>&1; tail -3 gpurun_out/r01m_pytest.log
timeout 300 python tools/bench_ops.py print > gpurun_out/r01m_print.json 2> gpurun_out/r01m_print.err; cat gpurun_out/r01m_print.json
timeout 600 python bench.py --ops > gpurun_out/r01m_bench_n1.json 2> gpurun_out/r01m_bench_n1.err; tail -2 gpurun_out/r01m_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r01m_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["chain"]["frac_of_aggregate_peak"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("cold",{}).get("value"), d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print(json.dumps(d.get("ops",{}))[:3000])
PY
timeout 600 python bench.py --steps 2 --warmup 3 --shards-limit 1 --no-cpu --no-cold --no-sweep > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mask_kernel|expand_kernel" -s 4 -c 2 -o gpurun_out/r01m_chain -f python bench.py --steps 2 --warmup 3 --shards-limit 1 --no-cpu --no-cold --no-sweep > gpurun_out/r01m_ncu1.log 2>&1; tail -1 gpurun_out/r01m_ncu1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-cold --no-sweep > gpurun_out/r01m_ncu2.log 2>&1; tail -1 gpurun_out/r01m_ncu2.log
