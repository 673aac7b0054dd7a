"""The composition sfb200_zgroup_factorize runs for a COMPLEX matrix on several GPUs -- partition of the embedded structure, host
expand / read-back -- on random matrices and 2-8 ranks: lock step and threaded protocol against the oracle (sf_oracle_z.c).
    python tests/tools/zpartition_fuzz.py [seed]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sforacle, sparseframe_b200 as sf
from test_irregular import random_matrix
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
f = sforacle._load_chk().sfcheck_factorize_complex_threaded
f.restype = C.c_long
f.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + [C.c_int] * 7 + [C.c_double, C.c_void_p]
for k in range(80):
    kind = str(rng.choice(["cholesky", "lu"]))
    if rng.random() < 0.5:
        gen, g = [("lap2d", int(rng.integers(10, 30))), ("lap3d", int(rng.integers(5, 11))), ("cd3d", int(rng.integers(5, 10)))][int(rng.integers(0, 3))]
        if kind == "cholesky" and gen == "cd3d": gen = "lap3d"
        n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky")); what = "%s %d" % (gen, g)
    else:
        n = int(rng.choice([100, 300, 700]))
        n, r, c, x, sym = random_matrix(n, float(rng.choice([0.004, 0.01, 0.03])) * 300 / n, int(rng.integers(1, 10**6)), kind == "cholesky", bool(rng.random() < 0.5))
        ph = rng.uniform(-0.6, 0.6, len(x))
        if kind == "cholesky": ph[r == c] = 0.0
        z = x * np.exp(1j * ph); what = "random %d" % n
    world = int(rng.choice([2, 3, 4, 8])); nb = int(rng.choice([0, 8, 16]))
    os.environ["SFCHECK_DIST_MIN_M"] = str(int(rng.choice([8, 32, 512]))); os.environ["SFCHECK_DIST_PANEL_MIN"] = str(int(rng.choice([16, 64, 2048])))
    print("case", k, kind, what, "world", world, "nb", nb, flush=True)
    A = sf.compress(n, r, c, z, sym)
    S = sf.analyze(A, kind)
    st = {q: S.array(q) for q in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    split = kind == "lu" and not sym
    if split: st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
    ref, bad = sforacle.zfactorize(kind, st, S.array("Lx"), S.array("Ux") if split else None)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    prm = dict(nb=nb, outer=2 * nb, small_a=0 if nb else -1, small_b=0 if nb else -1)
    out, bad = sforacle.check_plan_complex_partitioned(kind, n, A.Cp, A.Ci, A.Cx, sym, sf.DEFAULT_DEVSLOT, world, row_phase=bool(rng.random() < 0.5), xsize=S.int("xsize"), **prm)
    assert bad == 0 and np.abs(out[mask] - ref[mask]).max() <= 1e-11 * np.abs(ref).max()
    out2 = np.zeros(S.int("xsize"), dtype=np.complex128)
    Cz = np.ascontiguousarray(A.Cx, dtype=np.complex128)
    rc = f(int(kind == "lu"), int(sym), n, A.Cp.ctypes.data, A.Ci.ctypes.data, Cz.ctypes.data, sf.DEFAULT_DEVSLOT, prm["nb"], prm["outer"], prm["small_a"], prm["small_b"], world, 2, int(rng.choice([0, 20])), 60.0, out2.ctypes.data)
    assert rc == 0 and np.abs(out2[mask] - ref[mask]).max() <= 1e-11 * np.abs(ref).max()
print("complex partition fuzz: 80 cases equal to the oracle (lock step + threaded protocol)")
