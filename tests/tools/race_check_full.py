"""Happens-before race check (oracle/plan_race_check.cpp) of the FULL schedule of a BASELINE.json config, streamed copy-out
included.  Host only; cfg2 needs ~30 GB of shadow memory in lean mode and ~4 minutes.
    SFCHECK_RACE_LEAN=1 python tests/tools/race_check_full.py cfg2"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sparseframe_b200 as sf  # noqa: E402
import sforacle  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
gen, g, kind = sf.matrices.CONFIGS[cfg]
A = sf.compress(*sf.matrices.generate(gen, g))
t = time.time()
b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1)
print("%s: plan in %.1f s, %d launches, %.2f G elements (factor image + scratch)" % (
    cfg, time.time() - t, b.int("nlaunches"), b.int("xsize_work") / 1e9), flush=True)
fn = b.L.sfcheck_race_check
fn.restype = C.c_long
fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
t = time.time()
print("%s: %d unordered conflicting access pairs (main / side / diag / copy streams), %.1f s" % (cfg, fn(b.h, 0, 0), time.time() - t))
