"""Timing experiment on the largest GEMM-heavy launches only: one big front (dense SPD matrix as a single supernode chain)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
gen, g, kind = sf.matrices.CONFIGS[cfg]
S = sf.analyze(sf.compress(*sf.matrices.generate(gen, g)), kind)
lx = S.array("Lx")
for dbg in (0, 1, 2, 3):
    os.environ["SFB200_DBG"] = str(dbg)
    plan = sf.Plan(S)
    try:
        plan.factorize(lx, download=False)
    except RuntimeError:
        pass
    try:
        plan.factorize(lx, download=False)
    except RuntimeError:
        pass
    try:
        prof = plan.profile()
    except RuntimeError:
        prof = {}
    print("dbg", dbg, "numeric %.1f ms" % plan.last_ms(0), {k: "%.1fms" % v["ms"] for k, v in prof.items()}, flush=True)
    plan.close()
