"""cfg4 (LU, 3-D convection-diffusion 100^3) with the k-major in-panel updates on the classic cp.async kernel
(SFB200_WS_KMAJOR=0) and on the warp-specialized bulk-copy kernel (=1): device-resident numeric time, per-kernel split,
device residual.  One analysis, two plans."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import numpy as np
import sparseframe_b200 as sf

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
gen, g, kind = sf.matrices.CONFIGS[cfg]
t0 = time.time()
n, r, c, x, sym = sf.matrices.generate(gen, g)
A = sf.compress(n, r, c, x, sym)
S = sf.analyze(A, kind)
split = kind == "lu" and not sym
lx, ux = S.array("Lx"), (S.array("Ux") if split else None)
print("%s: n=%d nsuper=%d flops %.3e analyze %.1fs (order %.1f)" % (cfg, n, S.int("nsuper"), S.flops, time.time() - t0,
                                                                   S.double("t_order")), flush=True)
b = 1.0 + np.arange(n) / n
for flag in (0, 1, 0, 1):
    os.environ["SFB200_WS_KMAJOR"] = str(flag)
    plan = sf.Plan(S)
    ms = []
    for _ in range(4):
        plan.factorize(lx, ux, download=False)
        ms.append(plan.last_ms(0))
    xs = plan.solve(b)
    res = plan.residual(xs, b)
    prof = plan.profile()
    print("WS_KMAJOR=%d: numeric %s ms -> best %.2f ms = %.2f TFLOP/s | backward %.2e | launches %d" % (
        flag, ["%.1f" % m for m in ms], min(ms), S.flops / min(ms) / 1e9, res["rnorm"] / (res["anorm"] * res["xnorm"]),
        plan.int("nlaunches")), flush=True)
    print("    " + " ".join("%s %.1fms(%.1fTF)" % (k.replace("k_", ""), v["ms"], v["flops"] / v["ms"] / 1e9)
                            for k, v in prof.items()), flush=True)
    plan.close()
