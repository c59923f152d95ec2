#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep
(captured with --import-source on, built with -lineinfo).
python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, acc = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 10 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 10 and r[0].isdigit():
            ie = hdr.index("Instructions Executed")
            sm = hdr.index("# Samples")
            try:
                acc.append((int(r[ie]), int(r[sm]), fname, r[0], r[1]))
            except ValueError:
                pass
    tot = sum(a[0] for a in acc) or 1
    smp = sum(a[1] for a in acc) or 1
    print(f"total warp instructions {tot}, samples {smp}")
    for v, s, f, l, src in sorted(acc, reverse=True)[:top]:
        print(f"{100 * v / tot:5.1f}% inst {100 * s / smp:5.1f}% smp  {f}:{l}: {src.strip()[:110]}")


if __name__ == "__main__":
    main()
