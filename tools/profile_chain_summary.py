#!/usr/bin/env python
"""profiles/<tag>_chain_ncu_summary.md from the files tools/run_round_checks.sh leaves in
gpurun_out/: <tag>_chain.ncu-rep, <tag>_launches.csv, <tag>_bench_n1.json.
python tools/profile_chain_summary.py r01m "what changed"  """
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    g = os.path.join(ROOT, "gpurun_out")
    txt = open(os.path.join(g, f"{tag}_launches.csv")).read().splitlines()
    i = [k for k, line in enumerate(txt) if line.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("adb::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                         os.path.join(g, f"{tag}_chain.ncu-rep"), "--md"], stdout=subprocess.PIPE, text=True).stdout
    d = json.load(open(os.path.join(g, f"{tag}_bench_n1.json")))
    rf = d["roofline"]
    a_ms, b_ms = rf["avg_launch_ms"], rf["second_kernel"]["avg_launch_ms"]
    m = agg["mask_kernel"]
    e = next(v for k, v in agg.items() if k.startswith("expand_kernel<0, 1"))
    mu, eu = m[1] / 1e3 / m[0], e[1] / 1e3 / e[0]
    out = [f"# {tag}: the north-star chain as shipped (2 launches per shard), ncu --set full, one 500 M-row shard, 1 % selectivity", "",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:\"mask_kernel|expand_kernel\" -s 4 -c 2 "
           "python bench.py --steps 2 --warmup 3 --shards-limit 1 --no-cpu --no-cold --no-sweep` (after the same command exited 0 "
           f"without ncu).  Launch list of the full 8-shard step: `profiles/{tag}_launches.csv` (`ncu --metrics "
           "gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-cold "
           "--no-sweep`).  Both come from `tools/run_round_checks.sh`; this file from `tools/profile_chain_summary.py`.", "",
           md, "",
           "Launch list (all launches of the bench command, cold-cache serialised times; it also contains the e2e leg's "
           "operator calls: with the deferred select these are the same two kernels plus count_total / publish):", "", "| kernel | launches | total ms | avg us |", "|---|---|---|---|"]
    for k, (n, t) in agg.items():
        out.append(f"| {k} | {n} | {t / 1e6:.3f} | {t / 1e3 / n:.1f} |")
    out += ["", f"Per shard of the device-resident step: mask_kernel {mu:.1f} us + fused expansion {eu:.1f} us under ncu = "
            f"{100 * mu / (mu + eu):.1f} % / {100 * eu / (mu + eu):.1f} %; bench.py's live CUDA-event marks give {a_ms:.3f} ms / "
            f"{b_ms:.3f} ms = {100 * a_ms / (a_ms + b_ms):.1f} % / {100 * b_ms / (a_ms + b_ms):.1f} % "
            f"(profiles/{tag}_bench_n1.json): the shares agree.  Chain: {d['ms_per_step']:.3f} ms per 4 B rows = "
            f"{d['value'] / 1e12:.3f} T rows/s, {100 * d['chain']['frac_of_aggregate_peak']:.1f} % of the measured HBM peak on 4N + 20H."]
    if note:
        out += ["", note]
    with open(os.path.join(ROOT, "profiles", f"{tag}_chain_ncu_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    shutil.copy(os.path.join(g, f"{tag}_launches.csv"), os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    shutil.copy(os.path.join(g, f"{tag}_bench_n1.json"), os.path.join(ROOT, "profiles", f"{tag}_bench_n1.json"))
    print("\n".join(out))


if __name__ == "__main__":
    main()
