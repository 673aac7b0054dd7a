"""Per-launch efficiency table of one config: measured ms (CUDA events) against the FP64 tensor-pipe ideal."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
gen, g, kind = sf.matrices.CONFIGS[cfg]
S = sf.analyze(sf.compress(*sf.matrices.generate(gen, g)), kind)
plan = sf.Plan(S)
plan.factorize(S.array("Lx"), S.array("Ux") if (kind == "lu") else None, download=False)
plan.profile()
rows = plan.profile_launches()
peak = 37.0e12
names = {0: "smallA", 1: "smallB", 2: "potrf", 3: "getrf", 4: "trsm", 5: "gemm128", 6: "gemm64", 7: "gemmWS", 8: "trsm_mma"}
print("total profiled ms %.2f" % rows[:, 4].sum())
# per level summary
for lev in sorted(set(rows[:, 1].astype(int))):
    r = rows[rows[:, 1] == lev]
    line = "level %2d: %7.2f ms |" % (lev, r[:, 4].sum())
    for k in (5, 6, 2, 4, 0, 1):
        rk = r[r[:, 0] == k]
        if len(rk): line += " %s %d launches %.2f ms (%.1f TF/s)" % (names[k], len(rk), rk[:, 4].sum(), rk[:, 5].sum() / max(rk[:, 4].sum(), 1e-9) / 1e9)
    print(line)
# 64x64-tile kernel by K class
g6 = rows[rows[:, 0] == 6]
for lo, hi in ((0, 65), (65, 129), (129, 257), (257, 1025), (1025, 1e9)):
    sel = (g6[:, 6] >= lo) & (g6[:, 6] < hi)
    if sel.any():
        print("gemm64  Kmax in [%d,%d): %3d launches %7.2f ms  %.2f TF/s  ideal %.2f ms" % (lo, hi, sel.sum(), g6[sel, 4].sum(), g6[sel, 5].sum() / g6[sel, 4].sum() / 1e9, g6[sel, 5].sum() / peak * 1e3))
for k in (2, 4, 8):
    rk = rows[rows[:, 0] == k]
    if len(rk): print("%s: %d launches, %.2f ms, %.1f us per launch" % (names[k], len(rk), rk[:, 4].sum(), 1e3 * rk[:, 4].sum() / len(rk)))
# gemm128 launches by K class
g = rows[(rows[:, 0] == 5) | (rows[:, 0] == 7)]
for lo, hi in ((0, 65), (65, 257), (257, 1025), (1025, 1e9)):
    sel = (g[:, 6] >= lo) & (g[:, 6] < hi)
    if sel.any():
        print("gemm128 Kmax in [%d,%d): %3d launches %7.2f ms  %.2f TF/s  ideal %.2f ms" % (lo, hi, sel.sum(), g[sel, 4].sum(), g[sel, 5].sum() / g[sel, 4].sum() / 1e9, g[sel, 5].sum() / peak * 1e3))
print("slowest gemm128 launches: (level, grid, ops, ms, TF/s, kmax, kmin)")
for r in g[np.argsort(-g[:, 4])][:14]:
    print("   L%2d grid %6d ops %4d  %7.2f ms  %5.1f TF/s  K %5d..%5d" % (r[1], r[2], r[3], r[4], r[5] / r[4] / 1e9, r[7], r[6]))
