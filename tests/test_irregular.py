"""Irregular inputs (the PDE generators only produce grid-shaped elimination trees): random sparse SPD / unsymmetric
diagonally dominant matrices, forests (disconnected blocks), arrows, empty and 1 x 1 matrices.  Symbolic structure must be
bit-exact with the compiled reference, the interpreted schedule must reproduce its factor."""
import numpy as np
import pytest

import sforacle
import sfref
import sparseframe_b200 as sf

needs_ref = pytest.mark.skipif(not (sfref.available("cholesky") and sfref.available("lu")),
                               reason="compiled reference not built")
INT_ARRAYS = ("Perm", "Parent", "Post", "ColCount", "Lp", "Li", "Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "ST_Map",
              "ST_Pointer", "ST_Index", "LeafQueue")


def random_matrix(n, density, seed, symmetric, pattern_symmetric=True):
    """Triplets in a shuffled (file-like) order: lower triangle for symmetric, general otherwise; strictly diagonally
    dominant so that neither factorization needs pivoting."""
    rng = np.random.default_rng(seed)
    m = int(density * n * n / 2) + n
    i = rng.integers(0, n, m)
    j = rng.integers(0, n, m)
    keep = i > j
    i, j = i[keep], j[keep]
    key = np.unique(i * n + j)
    i, j = key // n, key % n
    v = -rng.random(len(i)) - 0.1
    if symmetric:
        rows, cols, vals = i, j, v
        rowsum = np.zeros(n)
        np.add.at(rowsum, i, np.abs(v))
        np.add.at(rowsum, j, np.abs(v))
    else:
        up = rng.random(len(i)) < (1.0 if pattern_symmetric else 0.6)          # some (j, i) partners are missing
        w = -rng.random(int(up.sum())) - 0.1
        rows = np.concatenate([i, j[up]])
        cols = np.concatenate([j, i[up]])
        vals = np.concatenate([v, w])
        rowsum = np.zeros(n)
        np.add.at(rowsum, rows, np.abs(vals))
        colsum = np.zeros(n)
        np.add.at(colsum, cols, np.abs(vals))
        rowsum = np.maximum(rowsum, colsum)
    rows = np.concatenate([rows, np.arange(n)])
    cols = np.concatenate([cols, np.arange(n)])
    vals = np.concatenate([vals, rowsum + 1.0])
    order = rng.permutation(len(rows))
    return n, rows[order].astype(np.int64), cols[order].astype(np.int64), vals[order], symmetric


def check_against_reference(n, r, c, x, sym, kind, slot=sf.DEFAULT_DEVSLOT, **prm):
    A = sf.compress(n, r, c, x, sym)
    ref = sfref.Reference(n, r, c, x, sym, kind=kind, devslot=slot).analyze().factorize()
    S = sf.analyze(A, kind, devslot=slot)
    for k in INT_ARRAYS:
        a, b = S.array(k), ref.array(k)
        if k == "LeafQueue":                     # the reference's factorize appends to it (CH.c:2976)
            a, b = a[:S.int("nsleaf")], b[:S.int("nsleaf")]
        assert np.array_equal(a, b[:len(a)]), k
    for k in ("nsuper", "isize", "xsize", "csize", "nstage", "nsleaf"):
        assert S.int(k) == ref.long(k), k
    out, bad, _ = sforacle.plan_check(kind, A.n, A.Cp, A.Ci, A.Cx, A.symmetric, slot, **prm)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, ref.array("Super"), ref.array("Lsip"), ref.array("Lsxp"), ref.long("xsize"))
    a, b = ref.array("Lsx")[mask], out[mask]
    assert sf.entrywise_rel_diff(a, b) <= 1e-10
    assert np.array_equal(a == 0, b == 0)
    ref.close()


@needs_ref
@pytest.mark.parametrize("n,density,seed", [(60, 0.08, 1), (300, 0.02, 2), (900, 0.006, 3), (1500, 0.002, 4)])
@pytest.mark.parametrize("prm", [dict(), dict(nb=8, outer=16, small_a=0, small_b=0)], ids=["default", "nb8"])
def test_random_spd_cholesky(n, density, seed, prm):
    check_against_reference(*random_matrix(n, density, seed, True), "cholesky", **prm)


@needs_ref
@pytest.mark.parametrize("n,density,seed,patsym", [(60, 0.08, 5, True), (300, 0.02, 6, False), (900, 0.006, 7, False),
                                                   (400, 0.01, 8, True)])
@pytest.mark.parametrize("prm", [dict(), dict(nb=8, outer=16, small_a=0, small_b=0)], ids=["default", "nb8"])
def test_random_unsymmetric_lu(n, density, seed, patsym, prm):
    check_against_reference(*random_matrix(n, density, seed, False, patsym), "lu", **prm)


@needs_ref
def test_random_symmetric_input_to_lu():
    check_against_reference(*random_matrix(250, 0.02, 9, True), "lu")


@needs_ref
@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_forest_of_disconnected_blocks(kind):
    """Three disconnected grids of different sizes plus isolated vertices: several etree roots."""
    parts, off = [], 0
    for gen, g in (("lap2d", 9), ("lap3d", 5), ("lap2d", 4)):
        n, r, c, x, _ = sf.matrices.generate(gen, g)
        parts.append((r + off, c + off, x))
        off += n
    iso = np.arange(off, off + 5)
    parts.append((iso, iso, np.full(5, 3.0)))
    n = off + 5
    r, c, x = (np.concatenate([p[k] for p in parts]) for k in range(3))
    check_against_reference(n, r, c, x, True, kind)


@needs_ref
def test_small_slot_caps_supernodes():
    """A device slot of a few KB forces the fundamental-supernode cap and refuses relaxed merges (CH.c:1476-1496,
    1556-1574): the structure depends on devSlotSize and must still match."""
    n, r, c, x, sym = sf.matrices.generate("lap3d", 7)
    for slot in (4096, 20000):
        check_against_reference(n, r, c, x, sym, "cholesky", slot=slot)


def test_empty_and_tiny_matrices():
    z = np.zeros(0, dtype=np.int64)
    S = sf.analyze(sf.compress(0, z, z, np.zeros(0), True), "cholesky")
    assert S.int("nsuper") == 0 and S.int("xsize") == 0 and S.flops == 0.0
    for kind in ("cholesky", "lu"):
        A = sf.compress(1, [0], [0], [4.0], True)
        out, bad, _ = sforacle.plan_check(kind, A.n, A.Cp, A.Ci, A.Cx, True, sf.DEFAULT_DEVSLOT)
        assert bad == 0 and out.tolist() == [2.0 if kind == "cholesky" else 4.0]
    # 2 x 2 unsymmetric
    A = sf.compress(2, [0, 1, 0, 1], [0, 0, 1, 1], [4.0, 2.0, 1.0, 3.0], False)
    out, bad, _ = sforacle.plan_check("lu", A.n, A.Cp, A.Ci, A.Cx, False, sf.DEFAULT_DEVSLOT)
    assert bad == 0
    S = sf.analyze(A, "lu")
    perm = S.array("Perm")
    dense = np.array([[4.0, 1.0], [2.0, 3.0]])[np.ix_(perm, perm)]
    l21 = dense[1, 0] / dense[0, 0]
    want = sorted([dense[0, 0], dense[0, 1], l21, dense[1, 1] - l21 * dense[0, 1]])
    assert np.allclose(sorted(out[out != 0]), want)


@pytest.mark.parametrize("n", [2, 3, 17, 70])
def test_dense_matrix_against_numpy(n):
    """A dense matrix: the first column's count is n.  The reference's weighted postorder drops such a column (bucket
    array of n entries, CH.c:1177-1200; out-of-bounds write, incomplete Post), so the check is against numpy here: one
    supernode chain, factor of P A P^T."""
    rng = np.random.default_rng(n)
    M = rng.random((n, n)) - 0.5
    spd = M @ M.T + n * np.eye(n)
    i, j = np.tril_indices(n)
    A = sf.compress(n, i, j, spd[i, j], True)
    S = sf.analyze(A, "cholesky")
    assert np.array_equal(np.sort(S.array("Post")), np.arange(n))
    perm = S.array("Perm")
    out, bad, _ = sforacle.plan_check("cholesky", A.n, A.Cp, A.Ci, A.Cx, True, sf.DEFAULT_DEVSLOT)
    assert bad == 0
    Lref = np.linalg.cholesky(spd[np.ix_(perm, perm)])
    Super, Lsip, Lsxp, Lsi = (S.array(k) for k in ("Super", "Lsip", "Lsxp", "Lsi"))
    got = np.zeros((n, n))
    for s in range(S.int("nsuper")):
        nc, nr = Super[s + 1] - Super[s], Lsip[s + 1] - Lsip[s]
        panel = out[Lsxp[s]:Lsxp[s] + nc * nr].reshape(nc, nr).T
        rows = Lsi[Lsip[s]:Lsip[s + 1]]
        got[np.ix_(rows, np.arange(Super[s], Super[s + 1]))] = panel
    assert np.allclose(np.tril(got), Lref, rtol=1e-11, atol=1e-13)
    # unsymmetric dense, LU without pivoting on a diagonally dominant matrix
    U = rng.random((n, n)) - 0.5 + n * np.eye(n)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    B = sf.compress(n, ii.ravel(), jj.ravel(), U.ravel(), False)
    S2 = sf.analyze(B, "lu")
    assert np.array_equal(np.sort(S2.array("Post")), np.arange(n))
    out2, bad2, _ = sforacle.plan_check("lu", B.n, B.Cp, B.Ci, B.Cx, False, sf.DEFAULT_DEVSLOT)
    assert bad2 == 0
    st = {k: S2.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li", "Up", "Ui")}
    st["csize"], st["nsleaf"] = S2.int("csize"), S2.int("nsleaf")
    b = 1.0 + np.arange(n) / n
    x = sforacle.solve("lu", st, out2, b)
    p2 = S2.array("Perm")
    assert np.allclose(U[np.ix_(p2, p2)] @ x, b, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("n,density,seed,kind,world,nb", [
    (300, 0.02, 11, "cholesky", 2, 0), (500, 0.006, 12, "cholesky", 5, 8), (400, 0.01, 13, "lu", 3, 0),
    (250, 0.03, 14, "lu", 8, 16), (600, 0.004, 15, "cholesky", 8, 0), (120, 0.1, 16, "lu", 4, 8)])
def test_partitioned_schedule_on_irregular_trees(n, density, seed, kind, world, nb):
    """Etree-subtree partition over `world` ranks on irregular (random) elimination trees and forests: all ranks simulated
    in one process, remote update entries and pushed blocks applied to their images; the union of the owned panels must be the single-rank factor."""
    import ctypes as C
    import os
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsfcheck.so"))
    lib.sfcheck_factorize_partitioned.restype = C.c_long
    lib.sfcheck_factorize_partitioned.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p] * 4
    nn, r, c, x, sym = random_matrix(n, density, seed, kind == "cholesky", seed % 2 == 0)
    A = sf.compress(nn, r, c, x, sym)
    S = sf.analyze(A, kind)
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    ux = None
    if kind == "lu":
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        ux = S.array("Ux")
    ref, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    out = np.full(S.int("xsize"), np.nan)
    bal = np.zeros(world)
    ncoll = (C.c_long * 2)()
    rc = lib.sfcheck_factorize_partitioned(int(kind == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data,
                                           A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT, nb, 2 * nb, 0 if nb else -1,
                                           0 if nb else -1, world, out.ctypes.data, bal.ctypes.data, ncoll, None)
    assert rc == 0
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    assert sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10
    assert bal.sum() == pytest.approx(S.flops, rel=1e-12)
    # the same partition with one free-running thread per rank and the real flags (groups of irregular forests overlap in
    # ways the regular grids never produce)
    lib.sfcheck_factorize_threaded.restype = C.c_long
    lib.sfcheck_factorize_threaded.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + \
        [C.c_int] * 7 + [C.c_double, C.c_void_p]
    for jitter in (0, 30):
        out2 = np.full(S.int("xsize"), np.nan)
        rc = lib.sfcheck_factorize_threaded(int(kind == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data,
                                            A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT, nb, 2 * nb, 0 if nb else -1, 0 if nb else -1,
                                            world, 2, jitter, 60.0, out2.ctypes.data)
        assert rc == 0
        assert sf.entrywise_rel_diff(ref[mask], out2[mask]) <= 1e-10
