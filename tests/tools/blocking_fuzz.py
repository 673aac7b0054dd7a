"""The CPU-interpreted schedule under random blockings (nb, outer, outer2, small-front thresholds, leading-dimension
alignment) on random matrices, against the oracle's restatement on the same structure: whatever the panel blocking, the
factor must be the same.
    python tests/tools/blocking_fuzz.py [cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sforacle  # noqa: E402
import sparseframe_b200 as sf  # noqa: E402
from test_irregular import random_matrix  # noqa: E402


def one_case(rng, k):
    kind = str(rng.choice(["cholesky", "lu"]))
    cplx = bool(rng.random() < 0.3)
    if rng.random() < 0.6:
        gen, g = [("lap2d", int(rng.integers(10, 60))), ("lap3d", int(rng.integers(5, 15))), ("cd3d", int(rng.integers(5, 13))),
                  ("el3d", int(rng.integers(2, 6)))][int(rng.integers(0, 4))]
        if kind == "cholesky" and gen == "cd3d":
            gen = "lap3d"
        if kind == "lu" and gen == "el3d":
            gen = "cd3d"
        if cplx:
            n, r, c, x, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
        else:
            n, r, c, x, sym = sf.matrices.generate(gen, g)
        what = "%s %d" % (gen, g)
    else:
        n = int(rng.choice([60, 300, 900]))
        n, r, c, x, sym = random_matrix(n, float(rng.choice([0.004, 0.01, 0.03])) * 300 / n, int(rng.integers(1, 10**6)),
                                        kind == "cholesky", bool(rng.random() < 0.5))
        if cplx:
            ph = rng.uniform(-0.6, 0.6, len(x))
            if kind == "cholesky":
                ph[r == c] = 0.0
            x = x * np.exp(1j * ph)
        what = "random n=%d" % n
    nb = int(rng.choice([8, 16, 32, 64]))              # the diagonal-block kernels take these orders
    outer = nb * int(rng.choice([1, 2, 4]))
    outer2 = outer * int(rng.choice([0, 2, 3, 4]))     # 0: the default third level (1024) stays
    prm = dict(nb=nb, outer=outer2 * 100000 + outer, small_a=int(rng.choice([-1, 0, 2048])), small_b=int(rng.choice([-1, 0, 8192])))
    if not cplx:
        prm["ld_align"] = int(rng.choice([0, 1, 2, 4, 16]))
    print("case %d: %s %s %s %s" % (k, kind, "complex" if cplx else "real", what, prm), flush=True)
    if os.environ.get("SFFUZZ_DUMP"):
        np.savez(os.environ["SFFUZZ_DUMP"], n=n, r=r, c=c, v=x, kind=kind, symmetric=sym, prm=str(prm))
    A = sf.compress(n, r, c, x, sym)
    S = sf.analyze(A, kind)
    st = {q: S.array(q) for q in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    split = kind == "lu" and not sym
    if split:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
    ref, bad = (sforacle.zfactorize if cplx else sforacle.factorize)(kind, st, S.array("Lx"), S.array("Ux") if split else None)
    assert bad == 0
    if cplx:
        out, bad, _, _ = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, sym, sf.DEFAULT_DEVSLOT,
                                                     row_phase=bool(rng.random() < 0.5), **prm)
    else:
        out, bad, _ = sforacle.plan_check(kind, n, A.Cp, A.Ci, A.Cx, sym, sf.DEFAULT_DEVSLOT, **prm)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    assert np.abs(out[mask] - ref[mask]).max() <= 1e-11 * np.abs(ref).max()
    assert np.array_equal(out[mask] == 0, ref[mask] == 0)


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    for k in range(cases):
        one_case(rng, k)
    print("blocking fuzz: %d schedules under random blockings equal to the oracle" % cases)
