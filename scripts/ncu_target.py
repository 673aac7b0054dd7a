"""Small fixed target for ncu: analyze + plan + ONE numeric factorization of a config (default cfg2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gen, g, kind = sf.matrices.CONFIGS[cfg] if cfg in sf.matrices.CONFIGS else (cfg.split(":")[0], int(cfg.split(":")[1]), cfg.split(":")[2])
n, r, c, x, sym = sf.matrices.generate(gen, g)
S = sf.analyze(sf.compress(n, r, c, x, sym), kind)
plan = sf.Plan(S)
split = kind == "lu" and not sym
for _ in range(reps):
    plan.factorize(S.array("Lx"), S.array("Ux") if split else None, download=False)
print(cfg, "launches", plan.last_launches(), "numeric ms %.3f" % plan.last_ms(0), "GFLOP/s %.1f" % (S.flops / plan.last_ms(0) / 1e6))
