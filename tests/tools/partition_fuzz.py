"""Randomised run of the multi-GPU host logic on the CPU (no GPU): etree partitions of random irregular matrices over 2-8
ranks, with the thresholds of the group-factorized panels and split Schur updates lowered so that small fronts take those
paths -- lock-step simulation (both rank orders) and the threaded, flag-accurate protocol check (oracle/plan_host_exec.cpp) --
against the single-rank oracle.
    python tests/tools/partition_fuzz.py [cases] [seed]"""
import ctypes as C
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

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsfcheck.so"))
lib.sfcheck_factorize_partitioned.restype = C.c_long
lib.sfcheck_factorize_partitioned.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p] * 4
lib.sfcheck_factorize_threaded.restype = C.c_long
lib.sfcheck_factorize_threaded.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + \
    [C.c_int] * 7 + [C.c_double, C.c_void_p]


def one_case(rng, k):
    kind = str(rng.choice(["cholesky", "lu"]))
    if rng.random() < 0.5:
        gen, g = [("lap2d", int(rng.integers(12, 40))), ("lap3d", int(rng.integers(6, 13))), ("cd3d", int(rng.integers(6, 12)))][int(rng.integers(0, 3))]
        if kind == "cholesky" and gen == "cd3d":
            gen = "lap3d"
        n, r, c, x, sym = sf.matrices.generate(gen, g)
        what = "%s %d" % (gen, g)
    else:
        n = int(rng.choice([150, 400, 900, 1500]))
        dens = float(rng.choice([0.004, 0.01, 0.03])) * 300 / n
        n, r, c, x, sym = random_matrix(n, dens, int(rng.integers(1, 10**6)), kind == "cholesky", bool(rng.random() < 0.5))
        what = "random n=%d dens=%.4f" % (n, dens)
    world = int(rng.choice([2, 3, 4, 5, 8]))
    nb = int(rng.choice([0, 8, 16]))
    os.environ["SFCHECK_DIST_MIN_M"] = str(int(rng.choice([8, 32, 512])))
    os.environ["SFCHECK_DIST_PANEL_MIN"] = str(int(rng.choice([16, 64, 2048])))
    print("case %d: %s %s world=%d nb=%d dist_min_m=%s panel_min=%s" % (k, kind, what, world, nb, os.environ["SFCHECK_DIST_MIN_M"],
                                                                     os.environ["SFCHECK_DIST_PANEL_MIN"]), flush=True)
    A = sf.compress(n, r, c, x, sym)
    S = sf.analyze(A, kind)
    st = {q: S.array(q) for q in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    ux = None
    if kind == "lu" and not sym:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        ux = S.array("Ux")
    ref, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    args = (int(kind == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data, A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT,
            nb, 2 * nb, 0 if nb else -1, 0 if nb else -1, world)
    for reverse in (False, True):
        if reverse:
            os.environ["SFCHECK_RANKS_REVERSED"] = "1"
        else:
            os.environ.pop("SFCHECK_RANKS_REVERSED", None)
        out = np.full(S.int("xsize"), np.nan)
        bal = np.zeros(world)
        ncoll = (C.c_long * 2)()
        rc = lib.sfcheck_factorize_partitioned(*args, out.ctypes.data, bal.ctypes.data, ncoll, None)
        assert rc == 0, rc
        assert sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10
        assert abs(bal.sum() - S.flops) <= 1e-9 * S.flops
    os.environ.pop("SFCHECK_RANKS_REVERSED", None)
    out2 = np.full(S.int("xsize"), np.nan)
    rc = lib.sfcheck_factorize_threaded(*args, 2, int(rng.choice([0, 20])), 60.0, out2.ctypes.data)
    assert rc == 0, rc
    assert sf.entrywise_rel_diff(ref[mask], out2[mask]) <= 1e-10
    return ncoll[0], ncoll[1]


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    tot = np.zeros(2, dtype=np.int64)
    for k in range(cases):
        tot += one_case(rng, k)
    print("partition fuzz: %d cases equal to the single-rank oracle (lock step in both rank orders + threaded protocol); "
          "%d barriers and %d pushes exercised" % (cases, tot[0], tot[1]))
