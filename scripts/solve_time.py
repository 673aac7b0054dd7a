"""Device solve timing (row f1): factorize once without copy-out, then solve a few times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import numpy as np
import sparseframe_b200 as sf
for cfg in sys.argv[1:]:
    gen, g, kind = sf.matrices.CONFIGS[cfg]
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    S = sf.analyze(sf.compress(n, r, c, x, sym), kind)
    split = kind == "lu" and not sym
    plan = sf.Plan(S)
    plan.factorize(S.array("Lx"), S.array("Ux") if split else None, download=False)
    b = 1.0 + np.arange(n) / n
    for it in range(4):
        xs = plan.solve(b)
        d = plan.residual(xs, b)
        print("%s solve %d: %.2f ms on the device, backward error %.2e" % (cfg, it, plan.last_ms(2), d["rnorm"] / (d["anorm"] * d["xnorm"])), flush=True)
    B = np.asfortranarray(np.stack([b] * 8, axis=1))
    plan.solve(B); plan.solve(B)
    print("%s solve with 8 right-hand sides: %.2f ms" % (cfg, plan.last_ms(2)), flush=True)
    plan.close()
