#!/usr/bin/env python
"""Generate the golden DSL fixtures from the reference's own test generators.

Runs /root/reference/project_tests/data_generation_scripts/milestone{1..4}.py UNMODIFIED
(imported from where they lie; nothing is copied) at a small table size and writes, per
test NN of the reference's suite (tests 1-37; 38-43 exercise update/delete, which the
reference's parser does not implement):

    tests/golden/dsl/testNNgen.dsl     the commands        (load paths use @GOLDEN@)
    tests/golden/dsl/testNNgen.exp     the expected output (computed by the generators with pandas)
    tests/golden/dsl/data*_generated.csv

The generators were written for pandas 1.x; two compatibility shims are installed at run
time: DataFrame.to_csv(line_terminator=) -> lineterminator=, and
DataFrame.append -> pandas.concat.

Usage (this container only -- /root/reference does not travel to the GPU box):
    python tests/golden/make_golden.py [rows=10000] [seed=42]
The committed fixtures were made with the defaults, which are the reference's own
(project_tests/data_generation_scripts/gen_all_for_staff_use.sh:9-14: 10000 rows, seed 42).
"""
import os
import runpy
import sys

import pandas as pd

REF_GEN = "/root/reference/project_tests/data_generation_scripts"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "dsl")
PLACEHOLDER = "@GOLDEN@"


def install_shims():
    orig_to_csv = pd.DataFrame.to_csv

    def to_csv(self, *a, **kw):
        if "line_terminator" in kw:
            kw["lineterminator"] = kw.pop("line_terminator")
        return orig_to_csv(self, *a, **kw)

    pd.DataFrame.to_csv = to_csv
    if not hasattr(pd.DataFrame, "append"):
        def append(self, other, ignore_index=False, **kw):
            if isinstance(other, (dict, pd.Series)):
                other = pd.DataFrame([other])
            return pd.concat([self, other], ignore_index=ignore_index)
        pd.DataFrame.append = append


def run(script, argv):
    old = sys.argv
    sys.argv = [script] + [str(a) for a in argv]
    try:
        runpy.run_path(os.path.join(REF_GEN, script), run_name="__main__")
    finally:
        sys.argv = old


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 42
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, REF_GEN)
    install_shims()
    cwd = os.getcwd()
    os.chdir(REF_GEN)          # the generators import data_gen_utils from their own directory
    try:
        run("milestone1.py", [rows, seed, OUT, PLACEHOLDER])
        run("milestone2.py", [rows, seed, OUT, PLACEHOLDER])
        run("milestone3.py", [rows, seed, OUT, PLACEHOLDER])
        # gen_all_for_staff_use.sh:9-14 -- fact, dim1, dim2 sizes, seed, zipf 1.0, 1000 distinct
        run("milestone4.py", [rows, rows, rows, seed, 1.0, max(10, rows // 10), OUT, PLACEHOLDER])
    finally:
        os.chdir(cwd)
    with open(os.path.join(OUT, "MANIFEST"), "w") as f:
        f.write(f"rows={rows} seed={seed} generator={REF_GEN} pandas={pd.__version__}\n")
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
