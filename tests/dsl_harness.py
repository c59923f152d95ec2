"""Replay of the reference's project_tests DSL files through a client/server pair, and the
reference verifier's comparison rules.  Test infrastructure only.

Mirrors /root/reference/infra_scripts/test_milestone.sh (server restarts before tests 2, 5,
11, 19, 20, 29, 32; one client process per test file, run_test.sh:20) and
infra_scripts/verify_output_standalone.sh (strip ANSI colours, `--` comments, blank lines
and surrounding whitespace; round any comma-separated field containing '.' to 2 decimals;
exact match first, then match after `sort -n` of both sides, :20-47).
"""
from __future__ import annotations

import os
import re
import shutil
import signal
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "dsl")
DROPIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "dropin")
RESTART_BEFORE = {2, 5, 11, 19, 20, 29, 32}          # test_milestone.sh:64
ANSI = re.compile(r"\x1B\[([0-9]{1,2}(;[0-9]{1,2})?)?[m|K]")


def clean(text: str) -> list[str]:
    """verify_output_standalone.sh:20."""
    out = []
    for line in text.splitlines():
        line = ANSI.sub("", line)
        line = re.sub(r"--.*$", "", line).strip()
        if not line:
            continue
        fields = []
        for f in line.split(","):
            if "." in f:
                try:
                    f = "%0.2f" % float(f)
                except ValueError:
                    pass
            fields.append(f)
        out.append(",".join(fields))
    return out


def _sort_n(lines):
    def key(s):
        m = re.match(r"\s*(-?\d+(\.\d+)?)", s)
        return (float(m.group(1)) if m else 0.0, s)
    return sorted(lines, key=key)


def verdict(out_text: str, exp_text: str) -> str:
    """'exact' | 'sorted' | 'fail' (verify_output_standalone.sh:31-47)."""
    a = clean(out_text)
    b = [" ".join(x.split()) for x in exp_text.splitlines() if x.strip()]
    a = [" ".join(x.split()) for x in a]
    if a == b:
        return "exact"
    if _sort_n(a) == _sort_n(b):
        return "sorted"
    return "fail"


def dsl_text(test_id: int) -> str:
    with open(os.path.join(GOLDEN, f"test{test_id:02d}gen.dsl")) as f:
        return f.read().replace("@GOLDEN@", GOLDEN)


def exp_text(test_id: int) -> str:
    with open(os.path.join(GOLDEN, f"test{test_id:02d}gen.exp")) as f:
        return f.read()


class ServerPair:
    """One server binary + its client (they share a compiled-in socket path)."""

    def __init__(self, flavour: str, workdir: str, env: dict | None = None):
        self.server = os.path.join(DROPIN, f"server_{flavour}")
        self.client = os.path.join(DROPIN, f"client_{flavour}")
        self.sock = f"/tmp/adb_{flavour}_unix_socket"
        self.workdir = workdir
        self.env = dict(os.environ, **(env or {}))
        self.proc = None
        self.server_log = os.path.join(workdir, f"server_{flavour}.log")
        os.makedirs(workdir, exist_ok=True)

    @staticmethod
    def available(flavour: str) -> bool:
        return all(os.path.exists(os.path.join(DROPIN, f"{x}_{flavour}")) for x in ("server", "client"))

    def alive(self) -> bool:
        return self.proc is not None and self.proc.poll() is None

    def start(self):
        self.stop()
        if os.path.exists(self.sock):
            os.unlink(self.sock)
        log = open(self.server_log, "ab")
        self.proc = subprocess.Popen([self.server], cwd=self.workdir, stdout=log, stderr=log,
                                     env=self.env, start_new_session=True)
        for _ in range(600):                       # CUDA context creation can take seconds
            if os.path.exists(self.sock):
                return
            if self.proc.poll() is not None:
                break
            time.sleep(0.05)
        raise RuntimeError(f"{self.server} did not come up; see {self.server_log}")

    def stop(self):
        if self.proc is None:
            return
        if self.proc.poll() is None:
            try:
                os.killpg(self.proc.pid, signal.SIGKILL)      # the exact group we started
            except ProcessLookupError:
                pass
            self.proc.wait()
        self.proc = None

    def run_dsl(self, text: str, timeout: float = 120.0) -> str:
        if not self.alive():
            self.start()
        r = subprocess.run([self.client], input=text.encode(), stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, cwd=self.workdir, timeout=timeout, env=self.env)
        # a `shutdown` command makes the server exit (or abort: the reference double-frees a
        # histogram on shutdown of indexed tables); give it a moment
        if re.search(r"^shutdown", text, flags=re.M):
            for _ in range(200):
                if not self.alive():
                    break
                time.sleep(0.025)
        return r.stdout.decode(errors="replace")

    def run_suite(self, test_ids) -> dict[int, str]:
        """Replay tests in order with the reference runner's restart points."""
        out = {}
        shutil.rmtree(os.path.join(self.workdir, "database"), ignore_errors=True)
        for t in test_ids:
            if t in RESTART_BEFORE or not self.alive():
                self.start()
            out[t] = self.run_dsl(dsl_text(t))
        self.stop()
        return out
