"""cfg4 (LU, 3-D convection-diffusion 100^3) with the in-panel updates' U11 operand (a) read row-major from the U11^T
scratch copy on the warp-specialized kernel (default), (b) k-major on the classic cp.async kernel (SFB200_LU_UT_COPY=0),
(c) k-major on the warp-specialized kernel, one 128-byte bulk copy per column and k-step (+ SFB200_WS_KMAJOR=1):
device-resident numeric time, per-kernel split, device residual.  One analysis, one plan per variant."""
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
variants = [("ut-copy", 1, 0), ("kmajor-classic", 0, 0), ("ut-copy", 1, 0), ("kmajor-classic", 0, 0)]
if "--with-kmajor-ws" in sys.argv:
    variants.append(("kmajor-ws", 0, 1))
for name, ut, flag in variants:
    os.environ["SFB200_LU_UT_COPY"] = str(ut)
    os.environ["SFB200_WS_KMAJOR"] = str(flag)
    plan = sf.Plan(S)
    ms = []
    for _ in range(4):
        plan.factorize(lx, ux, download=False)
        ms.append(plan.last_ms(0))
    xs = plan.solve(b)
    res = plan.residual(xs, b)
    prof = plan.profile()
    print("%s (scratch %.2f GB): numeric %s ms -> best %.2f ms = %.2f TFLOP/s | backward %.2e | launches %d" % (
        name, plan.int("scratch_elems") * 8e-9, ["%.1f" % m for m in ms], min(ms), S.flops / min(ms) / 1e9, res["rnorm"] / (res["anorm"] * res["xnorm"]),
        plan.int("nlaunches")), flush=True)
    print("    " + " ".join("%s %.1fms(%.1fTF)" % (k.replace("k_", ""), v["ms"], v["flops"] / v["ms"] / 1e9)
                            for k, v in prof.items()), flush=True)
    plan.close()
