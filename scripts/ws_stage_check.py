"""Factorize one generated matrix a few times and print the device backward error (kernel experiments)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import numpy as np
import sparseframe_b200 as sf
gen, g = sys.argv[1], int(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "cholesky"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n, r, c, x, sym = sf.matrices.generate(gen, g)
S = sf.analyze(sf.compress(n, r, c, x, sym), kind)
split = kind == "lu" and not sym
plan = sf.Plan(S)
b = 1.0 + np.arange(n) / n
for it in range(reps):
    plan.factorize(S.array("Lx"), S.array("Ux") if split else None, download=False)
    d = plan.residual(plan.solve(b), b)
    print("%s %d %s rep %d: %.2f ms, backward error %.2e" % (gen, g, kind, it, plan.last_ms(0), d["rnorm"] / (d["anorm"] * d["xnorm"])), flush=True)
