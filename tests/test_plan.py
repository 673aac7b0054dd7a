"""The engine's static schedule, interpreted on the CPU (oracle/plan_host_exec.cpp), against the golden factors and the
compiled reference: checks levels, blocked panel ops, ancestor segments, relative-index tables, padded layout, copy-out
without needing a GPU."""
import numpy as np
import pytest

import sforacle
import sfref
import sparseframe_b200 as sf
from conftest import build_matrix

PARAMS = [dict(), dict(small_a=0, small_b=0), dict(nb=16, outer=32, small_a=0, small_b=0),
          dict(nb=8, outer=16, small_b=3000), dict(ld_align=1), dict(ld_align=16, nb=32, outer=64),
          # three blocking levels (test hook of oracle/plan_host_exec.cpp: outer = outer2 * 100000 + outer)
          dict(nb=8, outer=64 * 100000 + 16, small_a=0, small_b=0), dict(nb=16, outer=128 * 100000 + 32, small_b=4096)]


@pytest.mark.parametrize("prm", PARAMS, ids=[str(p) for p in PARAMS])
def test_schedule_reproduces_golden(golden, prm):
    A, _ = build_matrix(golden["gen"], golden["grid"])
    out, bad, stats = sforacle.plan_check(golden["kind"], A.n, A.Cp, A.Ci, A.Cx, A.symmetric, golden["devslot"], **prm)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(golden["kind"], golden["Super"], golden["Lsip"], golden["Lsxp"], golden["xsize"])
    a, b = golden["Lsx"][mask], out[mask]
    assert sf.entrywise_rel_diff(a, b) <= 1e-10
    assert np.array_equal(a == 0, b == 0)


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind,prm", [
    ("lap3d", 20, "cholesky", dict()), ("lap2d", 100, "cholesky", dict(nb=8, outer=16, small_b=2000)),
    ("cd3d", 16, "lu", dict(nb=16, outer=32, small_a=0, small_b=0)), ("cd3d", 18, "lu", dict()),
])
def test_schedule_reproduces_compiled_reference(gen, g, kind, prm):
    A, (n, r, c, x, sym) = build_matrix(gen, g)
    ref = sfref.Reference(n, r, c, x, sym, kind=kind).analyze().factorize()
    out, bad, stats = sforacle.plan_check(kind, A.n, A.Cp, A.Ci, A.Cx, A.symmetric, sf.DEFAULT_DEVSLOT, **prm)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, ref.array("Super"), ref.array("Lsip"), ref.array("Lsxp"), ref.long("xsize"))
    a, b = ref.array("Lsx")[mask], out[mask]
    assert sf.entrywise_rel_diff(a, b) <= 1e-10
    assert np.array_equal(a == 0, b == 0)
    ref.close()


def test_schedule_flags_indefinite():
    n, r, c, x, sym = sf.matrices.generate("lap2d", 8)
    x = x.copy()
    x[r == c] = -4.0          # negative definite
    A = sf.compress(n, r, c, x, sym)
    _, bad, _ = sforacle.plan_check("cholesky", A.n, A.Cp, A.Ci, A.Cx, True, sf.DEFAULT_DEVSLOT)
    assert bad > 0


@pytest.mark.parametrize("gen,g,kind", [("lap3d", 14, "cholesky"), ("lap2d", 90, "cholesky"), ("cd3d", 12, "lu"), ("lap3d", 10, "lu"),
                                        ("el3d", 6, "cholesky")])
def test_solve_schedule_on_interpreter(gen, g, kind):
    """The device solve's task lists (64-column blocks, row slabs, LU's U11 / U12^T split), interpreted on the CPU:
    x must match the oracle's restatement of SparseFrame_solve_supernodal and give a tiny backward error."""
    import ctypes as C
    import os
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    backend = sforacle.CheckBackend(kind, A, sf.DEFAULT_DEVSLOT, 0, 1)
    lib = backend.L
    lib.sfcheck_factor_solve.restype = C.c_long
    lib.sfcheck_factor_solve.argtypes = [C.c_void_p, C.c_void_p]
    n = A.n
    b = 1.0 + np.arange(n) / n
    x = b.copy()
    assert lib.sfcheck_factor_solve(backend.h, x.ctypes.data) == 0
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    split = kind == "lu" and S.int("symmetric") == 0
    ux = None
    if split:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        ux = S.array("Ux")
    lsx, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    x_ref = sforacle.solve(kind, st, lsx, b)
    assert np.max(np.abs(x - x_ref)) <= 1e-11 * np.max(np.abs(x_ref))
    res, norms = sforacle.residual(kind, st, S.array("Lx"), b, x, ux)
    assert norms[3] / (norms[0] * norms[2]) <= 1e-12
    backend.close()


@pytest.mark.parametrize("gen,g,kind", [("cd3d", 16, "lu"), ("lap3d", 16, "cholesky"), ("lap3d", 12, "lu")])
def test_lookahead_side_launches_may_run_late(gen, g, kind, monkeypatch):
    """Look-ahead: the engine runs LF_SIDE launches on a second stream, any time up to the next LF_WAIT_SIDE launch.  The
    interpreter's default order is "immediately"; here they run as late as the contract allows (SFCHECK_SIDE_LATE), with
    a blocking small enough (8 / 16 / 32) that the top fronts are split.  Same factor either way, or the split
    regions overlap something the main stream does in between."""
    monkeypatch.setenv("SFCHECK_SIDE_LATE", "1")
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind)
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    split = kind == "lu" and S.int("symmetric") == 0
    ux = None
    if split:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        ux = S.array("Ux")
    ref, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    assert bad == 0
    out, bad, stats = sforacle.plan_check(kind, A.n, A.Cp, A.Ci, A.Cx, A.symmetric, sf.DEFAULT_DEVSLOT, nb=8,
                                          outer=32 * 100000 + 16, small_a=0, small_b=0)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    assert sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10


@pytest.mark.skipif(not (sfref.available("cholesky") and sfref.available("lu")), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind", [("lap2d", 14, "cholesky"), ("cd3d", 7, "lu"), ("lap3d", 6, "lu")])
def test_duplicate_entries_last_one_wins(gen, g, kind, monkeypatch):
    """The reader sums nothing (CH.c:526-587) and loadA assigns in storage order (CH.c:2018, LU.c:2498/2513): of several
    copies of one (i, j) the last stays.  The engine's assembly is one parallel scatter, so the plan switches the earlier
    copies off -- checked here against the compiled reference on a matrix where every 5th entry has a wrong-valued
    earlier twin.  (SFB200_METIS_KEEP_DUPLICATES: the reference hands METIS adjacency lists with repeated neighbours; the
    engine removes them by default -- see test_heavily_duplicated_entries_do_not_reach_metis -- and this switch reproduces the
    reference's ordering for the comparison.)"""
    monkeypatch.setenv("SFB200_METIS_KEEP_DUPLICATES", "1")
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    twin = np.arange(0, len(r), 5)
    r2 = np.concatenate([r[twin], r])            # the wrong copies come first in file order
    c2 = np.concatenate([c[twin], c])
    x2 = np.concatenate([x[twin] * 3.0 + 1.0, x])
    A = sf.compress(n, r2, c2, x2, sym)
    ref = sfref.Reference(n, r2, c2, x2, sym, kind=kind).analyze().factorize()
    out, bad, stats = sforacle.plan_check(kind, A.n, A.Cp, A.Ci, A.Cx, A.symmetric, sf.DEFAULT_DEVSLOT)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, ref.array("Super"), ref.array("Lsip"), ref.array("Lsxp"), ref.long("xsize"))
    a, b = ref.array("Lsx")[mask], out[mask]
    assert sf.entrywise_rel_diff(a, b) <= 1e-10
    ref.close()


@pytest.mark.parametrize("gen,g,kind", [("lap2d", 4, "cholesky"), ("lap2d", 9, "cholesky"), ("cd3d", 4, "lu"), ("lap3d", 5, "lu")])
def test_heavily_duplicated_entries_do_not_reach_metis(gen, g, kind):
    """Every entry four times (three wrong-valued copies first): the reference builds METIS adjacency lists with repeated
    neighbours (CH.c:920-931), which METIS does not allow -- on inputs like these it crashes, hangs or returns something that
    is not a permutation in one case out of four (the reference with it).  The engine removes repeated neighbours first: the
    ordering is a permutation, and the factor is the one of the matrix made of the LAST copies (oracle restatement on the
    engine's own structure)."""
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    r4 = np.concatenate([r, r, r, r])
    c4 = np.concatenate([c, c, c, c])
    x4 = np.concatenate([x * 3.0 + 1.0, -x, x * 0.5, x])
    A = sf.compress(n, r4, c4, x4, sym)
    S = sf.analyze(A, kind)
    assert np.array_equal(np.sort(S.array("Perm")), np.arange(n))
    out, bad, _ = sforacle.plan_check(kind, n, A.Cp, A.Ci, A.Cx, sym, sf.DEFAULT_DEVSLOT)
    assert bad == 0
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    ux = None
    if kind == "lu" and not sym:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        ux = S.array("Ux")
    ref, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], S.int("xsize"))
    assert bad == 0 and sf.entrywise_rel_diff(ref[mask], out[mask]) <= 1e-10
    # ... and that is the factor of the matrix without the wrong copies, analysed with the same ordering
    Sc = sf.analyze(sf.compress(n, r, c, x, sym), kind, order=S.array("Order"))
    for k in ("Perm", "Super", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(Sc.array(k), S.array(k)), k
    stc = {k: Sc.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    stc["csize"], stc["nsleaf"] = Sc.int("csize"), Sc.int("nsleaf")
    uxc = None
    if kind == "lu" and not sym:
        stc["Up"], stc["Ui"] = Sc.array("Up"), Sc.array("Ui")
        uxc = Sc.array("Ux")
    clean, bad = sforacle.factorize(kind, stc, Sc.array("Lx"), uxc)
    assert bad == 0 and sf.entrywise_rel_diff(clean[mask], out[mask]) <= 1e-10
    _, backward = sforacle.validate(kind, stc, Sc.array("Lx"), out, uxc)
    assert backward <= 1e-13
