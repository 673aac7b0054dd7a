"""Happens-before race check (oracle/plan_race_check.cpp) of the schedules of random irregular matrices under random blockings
-- three blocking levels with both look-aheads, narrow blocks, everything-large -- on one rank and on every rank of random
partitions.  Expected: zero hazards everywhere.
    python tests/tools/race_fuzz.py [cases] [seed]"""
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


def race_check(backend, by_segment):
    fn = backend.L.sfcheck_race_check
    fn.restype = C.c_long
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return fn(backend.h, by_segment, 0)


def one_case(rng, k):
    kind = str(rng.choice(["cholesky", "lu"]))
    if rng.random() < 0.6:
        gen, g = [("lap2d", int(rng.integers(20, 90))), ("lap3d", int(rng.integers(8, 22))), ("cd3d", int(rng.integers(8, 18))),
                  ("el3d", int(rng.integers(3, 8)))][int(rng.integers(0, 4))]
        if kind == "cholesky" and gen == "cd3d":
            gen = "lap3d"
        if kind == "lu" and gen == "el3d":
            gen = "cd3d"
        n, r, c, x, sym = sf.matrices.generate(gen, g)
        what = "%s %d" % (gen, g)
    else:
        n = int(rng.choice([300, 900, 2000]))
        dens = float(rng.choice([0.004, 0.01, 0.03])) * 300 / n
        n, r, c, x, sym = random_matrix(n, dens, int(rng.integers(1, 10**6)), kind == "cholesky", bool(rng.random() < 0.5))
        what = "random n=%d" % n
    nb = int(rng.choice([8, 16, 32, 64]))
    outer = nb * int(rng.choice([2, 4]))
    outer2 = outer * int(rng.choice([0, 2, 4]))
    prm = dict(nb=nb, outer=outer2 * 100000 + outer)
    if rng.random() < 0.5:
        prm.update(small_a=0, small_b=0)
    world = int(rng.choice([1, 1, 2, 3, 4, 8]))
    os.environ["SFCHECK_DIST_MIN_M"] = str(int(rng.choice([16, 64, 512])))
    os.environ["SFCHECK_DIST_PANEL_MIN"] = str(int(rng.choice([32, 128, 2048])))
    print("case %d: %s %s %s world=%d" % (k, kind, what, prm, world), flush=True)
    A = sf.compress(n, r, c, x, sym)
    hazards = 0
    for rank in range(world):
        b = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, rank, world, **prm)
        hazards += race_check(b, 1 if world > 1 else 0)
        b.close()
    assert hazards == 0, hazards
    return world


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    ranks = sum(one_case(rng, k) for k in range(cases))
    print("race fuzz: %d schedules (%d rank schedules) race-free" % (cases, ranks))
