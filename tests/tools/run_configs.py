"""All single-GPU BASELINE.json configs: numeric time, GFLOP/s, backward error (oracle solve + residual), per-kernel split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import sparseframe_b200 as sf
import sforacle
for cfg in sys.argv[1:]:
    gen, g, kind = sf.matrices.CONFIGS[cfg]
    t0 = time.time()
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    A = sf.compress(n, r, c, x, sym)
    S = sf.analyze(A, kind)
    ta = time.time() - t0
    split = kind == "lu" and not sym
    plan = sf.Plan(S)
    lx, ux = S.array("Lx"), (S.array("Ux") if split else None)
    pinned = torch.empty(S.int("xsize"), dtype=torch.float64).pin_memory()
    best = 1e30
    for _ in range(3):
        lsx = plan.factorize(lx, ux, out=pinned.numpy())
        best = min(best, plan.last_ms(0))
    e2e = plan.last_ms(1)
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    if split:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
    res, bwd = sforacle.validate(kind, st, lx, lsx, ux)
    prof = plan.profile()
    print("%s %s g=%d %s: n=%d nnz=%d nsuper=%d factor %.2f GB flops %.3e | analyze %.1fs | numeric %.2f ms = %.2f TFLOP/s, e2e %.2f ms | "
          "residual %.2e backward %.2e | launches %d" % (cfg, gen, g, kind, n, A.Cp[-1], S.int("nsuper"), S.int("xsize") * 8e-9, S.flops, ta,
                                                         best, S.flops / best / 1e9, e2e, res, bwd, plan.int("nlaunches")), flush=True)
    print("    " + " ".join("%s %.1fms(%.1fTF)" % (k.replace("k_", ""), v["ms"], v["flops"] / v["ms"] / 1e9) for k, v in prof.items()), flush=True)
    plan.close(); S.close(); del pinned
