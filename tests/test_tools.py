"""The randomised host-side checks under tests/tools/ (differential run against the compiled reference, partitions, races,
blockings, solves; logs of the long runs are under profiles/r02_*_fuzz.log) -- a few cases of each, so that the tools stay
runnable."""
import os
import subprocess
import sys

import pytest

import sfref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tests", "tools")


def run_tool(name, *args):
    p = subprocess.run([sys.executable, os.path.join(TOOLS, name)] + [str(a) for a in args], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@pytest.mark.skipif(not (sfref.available("cholesky") and sfref.available("lu")), reason="compiled reference not built")
def test_differential_fuzz_runs():
    out = run_tool("differential_fuzz.py", 12, 21)
    assert "differential fuzz:" in out and " 0 cases identical" not in out


def test_partition_fuzz_runs():
    assert "partition fuzz: 4 cases equal" in run_tool("partition_fuzz.py", 4, 21)


def test_race_fuzz_runs():
    assert "race fuzz: 4 schedules" in run_tool("race_fuzz.py", 4, 21)


def test_blocking_fuzz_runs():
    assert "blocking fuzz: 6 schedules" in run_tool("blocking_fuzz.py", 6, 21)
