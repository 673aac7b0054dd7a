"""Complex factorization on one GPU (row f4): timing of sfb200_zfactorize_device on the real embedding, with the compiled
reference's zherk/zgemm/zpotrf (or zgemm/zgetrf_nopiv) CPU path timed beside it and a parity check against it.
Usage: python scripts/zbench.py [gen grid kind]...   (default: lap3d 40 cholesky, cd3d 32 lu).  Prints one JSON line per case.
"complex_gflops" counts the flops a native complex kernel needs (4 real multiply-adds per complex one = 4x the real count of
the same structure); the embedding executes twice that ("executed_gflops")."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sparseframe_b200 as sf  # noqa: E402


def run(gen, g, kind, steps=5, with_ref=True):
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    A = sf.compress(n, r, c, z, sym)
    t0 = time.time()
    S = sf.analyze(A, kind)
    t_an = time.time() - t0
    lz = S.array("Lx")
    uz = S.array("Ux") if (kind == "lu" and not sym) else None
    t0 = time.time()
    plan = sf.ZPlan(S)
    t_plan = time.time() - t0
    ms = []
    for _ in range(steps):
        plan.factorize(lz, uz, download=False)
        ms.append(plan.last_ms())
    real_flops = S.flops                       # real flop count of the complex structure's shape (SURVEY 8d formulas)
    best = min(ms[1:]) if len(ms) > 1 else ms[0]
    b = (1 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n))
    x = plan.solve(b)
    res = plan.residual(x, b)["residual"]
    t0 = time.time()
    out = plan.factorize(lz, uz) if with_ref else None        # host buffers in, complex factor out (pageable memory here)
    t_e2e = time.time() - t0 if with_ref else None
    line = dict(case="%s %d %s complex" % (gen, g, kind), n=n, nsuper=S.int("nsuper"), xsize_complex=S.int("xsize"),
                analyze_s=round(t_an, 3), plan_s=round(t_plan, 3), ms_all=[round(v, 3) for v in ms], ms=round(best, 3),
                complex_gflops=round(4 * real_flops / best / 1e6, 1), executed_gflops=round(plan.real.double("flops") / best / 1e6, 1),
                launches=plan.real.last_launches(), e2e_host_s=None if t_e2e is None else round(t_e2e, 3), residual=res,
                embedded_factor_GB=round(8e-9 * plan.real.int("xsize_padded"), 3))
    if with_ref:
        import sfref
        if sfref.available(kind):
            R = sfref.Reference(n, r, c, z, sym, kind).analyze()
            t0 = time.time()
            R.factorize()
            t_ref = time.time() - t0
            ref = R.array("Lsx")
            st = R.structure()
            mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], R.long("xsize"))
            line.update(cpu_reference_s=round(t_ref, 3), cpu_reference_complex_gflops=round(4 * real_flops / t_ref / 1e9, 1),
                        max_rel_diff_vs_reference=sf.entrywise_rel_diff(out[mask], ref[mask]))
            R.close()
    plan.close()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    no_ref = "--no-ref" in args             # skip the compiled reference (minutes of CPU at the BASELINE sizes)
    procs = 1
    for a in list(args):
        if a.startswith("--order-procs="):  # opt-in parallel nested dissection (not the reference's permutation)
            procs = int(a.split("=")[1])
    args = [a for a in args if not a.startswith("--")]
    if procs > 1:
        from sparseframe_b200 import _capi
        _capi.lib().sfb200_set_ordering_processes(procs)
    cases = [(args[i], int(args[i + 1]), args[i + 2]) for i in range(0, len(args) - 2, 3)] or [("lap3d", 40, "cholesky"), ("cd3d", 32, "lu")]
    for gen, g, kind in cases:
        run(gen, g, kind, steps=3 if no_ref else 5, with_ref=not no_ref)
