"""Timing sweep over plan parameters for one config: analysis once, then one plan per environment setting."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
gen, g, kind = sf.matrices.CONFIGS[cfg]
n, r, c, x, sym = sf.matrices.generate(gen, g)
S = sf.analyze(sf.compress(n, r, c, x, sym), kind)
split = kind == "lu" and not sym
lx, ux = S.array("Lx"), (S.array("Ux") if split else None)
KEYS = ("SFB200_NB", "SFB200_OUTER", "SFB200_SMALL_A", "SFB200_SMALL_B", "SFB200_GEMM128_MIN", "SFB200_GEMM_WS", "SFB200_OUTER2", "SFB200_LOOKAHEAD", "SFB200_SPARE_SMS", "SFB200_LEAF_LOOKAHEAD", "SFB200_LEAVE_SMS", "SFB200_GEMM128_MIN_K", "SFB200_LU_UT_COPY", "SFB200_WS_KMAJOR")
envs = [dict()]
for spec in sys.argv[2:]:
    envs.append(dict(kv.split("=") for kv in spec.split(",")))
for env in envs:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    plan = sf.Plan(S)
    best = 1e30
    for _ in range(3):
        plan.factorize(lx, ux, download=False)
        best = min(best, plan.last_ms(0))
    prof = plan.profile()
    line = "%s env=%s launches=%d numeric %.2f ms  %.2f TFLOP/s |" % (cfg, env, plan.int("nlaunches"), best, S.flops / best / 1e9)
    for k, v in prof.items():
        line += " %s %.1fms(%d, %.1fTF)" % (k.replace("k_", ""), v["ms"], v["launches"], v["flops"] / v["ms"] / 1e9)
    if os.environ.get("SFB200_SWEEP_VALIDATE", "1") != "0":       # device solve + residual for every setting
        import numpy as np
        b = 1.0 + np.arange(n) / n
        d = plan.residual(plan.solve(b), b)
        line += " | device backward error %.2e" % (d["rnorm"] / (d["anorm"] * d["xnorm"]))
    print(line, flush=True)
    if not env and os.environ.get("SFB200_SWEEP_SOLVE"):       # device solve timing, 1 and 8 right-hand sides
        import numpy as np
        b = 1.0 + np.arange(n) / n
        for nrhs in (1, 8):
            B = b if nrhs == 1 else np.asfortranarray(np.stack([b * (k + 1) for k in range(nrhs)], axis=1))
            ms = []
            for _ in range(4):
                X = plan.solve(B)
                ms.append(plan.last_ms(2))
            xlast = X if nrhs == 1 else X[:, nrhs - 1] / nrhs
            d = plan.residual(np.ascontiguousarray(xlast), b)
            print("   solve nrhs=%d: %s ms, backward error of the last column %.2e" % (
                nrhs, ["%.2f" % m for m in ms], d["rnorm"] / (d["anorm"] * d["xnorm"])), flush=True)
    plan.close()
