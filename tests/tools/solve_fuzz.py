"""The device solve's task lists (forward / backward sweeps) interpreted on the CPU for random matrices and right-hand sides, against the
oracle's restatement of SparseFrame_solve_supernodal.   python tests/tools/solve_fuzz.py [seed]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sforacle, sparseframe_b200 as sf
from test_irregular import random_matrix
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for k in range(120):
    kind = str(rng.choice(["cholesky", "lu"]))
    if rng.random() < 0.5:
        gen, g = [("lap2d", int(rng.integers(10, 70))), ("lap3d", int(rng.integers(5, 16))), ("cd3d", int(rng.integers(5, 13)))][int(rng.integers(0, 3))]
        if kind == "cholesky" and gen == "cd3d": gen = "lap3d"
        n, r, c, x, sym = sf.matrices.generate(gen, g); what = "%s %d" % (gen, g)
    else:
        n = int(rng.choice([60, 300, 900, 2000]))
        n, r, c, x, sym = random_matrix(n, float(rng.choice([0.004, 0.01, 0.03])) * 300 / n, int(rng.integers(1, 10**6)), kind == "cholesky", bool(rng.random() < 0.5)); what = "random %d" % n
    print("case", k, kind, what, flush=True)
    A = sf.compress(n, r, c, x, sym)
    S = sf.analyze(A, kind)
    backend = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1)
    lib = backend.L
    lib.sfcheck_factor_solve.restype = C.c_long
    lib.sfcheck_factor_solve.argtypes = [C.c_void_p, C.c_void_p]
    b = rng.standard_normal(n)
    xx = b.copy()
    assert lib.sfcheck_factor_solve(backend.h, xx.ctypes.data) == 0
    st = {q: S.array(q) for q in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    split = kind == "lu" and not sym
    ux = None
    if split:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui"); ux = S.array("Ux")
    lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    x_ref = sforacle.solve(kind, st, lsx, b)
    assert np.max(np.abs(xx - x_ref)) <= 1e-10 * np.max(np.abs(x_ref)), k
    backend.close()
print("solve fuzz: 120 solve schedules equal to the oracle's substitution")
