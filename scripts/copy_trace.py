"""Timeline of the streamed copy-out (SFB200_COPY_TRACE): per etree level, when its kernels and when its copies ended."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import numpy as np
import torch
import sparseframe_b200 as sf
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
gen, g, kind = sf.matrices.CONFIGS[cfg]
n, r, c, x, sym = sf.matrices.generate(gen, g)
S = sf.analyze(sf.compress(n, r, c, x, sym), kind)
split = kind == "lu" and not sym
plan = sf.Plan(S)
lx, ux = S.array("Lx"), (S.array("Ux") if split else None)
pinned = torch.empty(S.int("xsize"), dtype=torch.float64).pin_memory()
for it in range(4):
    if it == 3:
        os.environ["SFB200_COPY_TRACE"] = "1"
    plan.factorize(lx, ux, out=pinned.numpy())
    print("iter %d: numeric %.2f ms, e2e %.2f ms" % (it, plan.last_ms(0), plan.last_ms(1)), flush=True)
Lsxp, Super = S.array("Lsxp"), S.array("Super")
print("factor %.2f GB" % (S.int("xsize") * 8e-9))
