"""Join an ncu launch list of ONE factorization (scripts/ncu_target.py, --metrics gpu__time_duration.sum) with the plan of
the same config built on the host: per launch kind, level, K range, algorithmic flops and the (cold-cache, serialised) ncu
duration; then totals per kernel and per K class.  Host only.
    python tests/tools/launch_table_from_ncu.py cfg2 profiles/r01_launches_cfg2_v7.csv"""
import collections
import csv
import ctypes as C
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import sparseframe_b200 as sf  # noqa: E402
import sforacle  # noqa: E402

cfg, path = sys.argv[1], sys.argv[2]
gen, g, kind = sf.matrices.CONFIGS[cfg]
A = sf.compress(*sf.matrices.generate(gen, g))
b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1)
fn = b.L.sfcheck_launch_info
fn.restype = C.c_long
fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
plan = []
for i in range(b.int("nlaunches")):
    out = np.zeros(8)
    assert fn(b.h, i, out.ctypes.data) == 0
    plan.append(out)
KNAME = {0: "k_small", 1: "k_small", 2: "k_diag", 3: "k_diag", 4: "k_trsm", 5: "k_gemm<", 6: "k_gemm<", 7: "k_gemm_ws"}
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
rows = [r for r in rows if "k_assemble" not in r[4]]
assert len(rows) == len(plan), (len(rows), len(plan))
cls = collections.OrderedDict()
detail = []
for r, p in zip(rows, plan):
    name = re.sub(r"\(.*", "", r[4])
    assert KNAME[int(p[0])] in name, (name, p[0])
    grid = int(re.sub(r"[^0-9,]", "", r[8]).split(",")[0])
    assert grid <= int(p[2]) or int(p[0]) != 7, (grid, p[2])         # the persistent kernel launches min(tiles, SMs) CTAs
    ms = float(r[14]) / 1e6
    k = int(p[0])
    if k in (5, 6, 7):
        kc = "K<=64" if p[6] <= 64 else "K<=256" if p[6] <= 256 else "K<=1024" if p[6] <= 1024 else "K>1024"
        key = ("gemm_ws 128x128" if k == 7 else "gemm 64x64" if k == 6 else "gemm 128x128") + " " + kc
    else:
        key = {0: "small", 1: "small", 2: "diag", 3: "diag", 4: "trsm"}[k]
    detail.append((ms, key, int(p[1]), int(p[2]), int(p[3]), int(p[7]), int(p[6]), p[5], int(p[4])))
    c = cls.setdefault(key, [0, 0.0, 0.0])
    c[0] += 1; c[1] += ms; c[2] += p[5]
total = sum(c[1] for c in cls.values())
print("%s: %d launches, %.1f ms of kernel time in the ncu list (serialised, cold cache)" % (cfg, len(plan), total))
for key, (n, ms, fl) in sorted(cls.items()):
    print("  %-26s %5d launches %9.2f ms %5.1f %%  %7.2f TFLOP/s" % (key, n, ms, 100 * ms / total, fl / ms / 1e9))
print("slowest launches: ms, class, level, tiles (CTAs for non-persistent kernels), ops, K range, TFLOP/s, waves of 148 SMs, launch flags")
for ms, key, lev, grid, ops, kmin, kmax, fl, flags in sorted(detail, reverse=True)[:28]:
    per_wave = 148 * (3 if key.startswith("gemm 64") else 1)
    print("  %8.3f  %-24s L%-2d %6d tiles %3d ops  K %5d..%5d  %6.2f TFLOP/s  %6.2f waves  flags %d" % (
        ms, key, lev, grid, ops, kmin, kmax, fl / ms / 1e9, grid / per_wave, flags))
