"""Host symbolic analysis (csrc/symbolic.cpp) must be BIT-EXACT with the reference's SparseFrame_analyze."""
import numpy as np
import pytest

import sfref
import sparseframe_b200 as sf
from conftest import build_matrix

INT_ARRAYS = ["Perm", "Parent", "Post", "ColCount", "Lp", "Li", "LTp", "LTi", "Super", "SuperMap", "Sparent", "Lsip",
              "Lsxp", "Lsi", "ST_Map", "ST_Pointer", "ST_Index", "LeafQueue", "Aoffset", "Moffset"]


def test_structure_bit_exact_with_golden(golden):
    A, _ = build_matrix(golden["gen"], golden["grid"])
    S = sf.analyze(A, golden["kind"], devslot=golden["devslot"])
    for k in ("nsuper", "isize", "xsize", "csize", "nstage", "nsleaf"):
        assert S.int(k) == golden[k], k
    names = list(INT_ARRAYS)
    if "Up" in golden:
        names += ["Up", "Ui", "UTp", "UTi"]
    for k in names:
        assert np.array_equal(S.array(k), golden[k]), k
    assert np.array_equal(S.array("Lx"), golden["Lx"])
    if "Ux" in golden:
        assert np.array_equal(S.array("Ux"), golden["Ux"])


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind,slot", [
    ("lap2d", 60, "cholesky", sf.DEFAULT_DEVSLOT), ("lap3d", 12, "cholesky", sf.DEFAULT_DEVSLOT),
    ("lap2d", 150, "cholesky", 200_000), ("lap3d", 30, "cholesky", 3_000_000), ("el3d", 6, "cholesky", sf.DEFAULT_DEVSLOT),
    ("cd3d", 12, "lu", sf.DEFAULT_DEVSLOT), ("cd3d", 20, "lu", 1_000_000), ("lap3d", 16, "lu", sf.DEFAULT_DEVSLOT),
])
def test_structure_bit_exact_with_compiled_reference(gen, g, kind, slot):
    A, (n, r, c, x, sym) = build_matrix(gen, g)
    ref = sfref.Reference(n, r, c, x, sym, kind=kind, devslot=slot).analyze()
    assert np.array_equal(A.Cp, ref.array("Cp")) and np.array_equal(A.Ci, ref.array("Ci"))
    S = sf.analyze(A, kind, devslot=slot)
    names = list(INT_ARRAYS) + (["Up", "Ui", "UTp", "UTi"] if kind == "lu" and not sym else [])
    for k in names:
        a, b = S.array(k), ref.array(k)
        assert np.array_equal(a, b[:len(a)]), k
    for k in ("nsuper", "isize", "xsize", "csize", "nstage", "nsleaf"):
        assert S.int(k) == ref.long(k), k
    assert np.array_equal(S.array("Lx"), ref.array("Lx"))
    ref.close()


def test_diagonal_matrix_identity_order():
    # no off-diagonals: METIS is skipped and the order is the identity (CH.c:933-939)
    n = 7
    A = sf.compress(n, np.arange(n), np.arange(n), np.arange(1.0, n + 1), True)
    S = sf.analyze(A, "cholesky")
    assert np.array_equal(np.sort(S.array("Perm")), np.arange(n))
    assert S.int("xsize") == S.array("Lsxp")[-1]
    assert np.all(S.array("Parent") == -1)


def test_one_by_one():
    A = sf.compress(1, [0], [0], [4.0], True)
    S = sf.analyze(A, "cholesky")
    assert S.int("nsuper") == 1 and S.int("xsize") == 1 and S.array("Lsi").tolist() == [0]


def test_flop_formula():
    A, _ = build_matrix("lap3d", 10)
    S = sf.analyze(A, "cholesky")
    nc = np.diff(S.array("Super")).astype(float)
    m = np.diff(S.array("Lsip")) - nc
    assert S.flops == pytest.approx(float((nc ** 3 / 3 + m * nc ** 2 + m * (m + 1) * nc).sum()), rel=1e-12)


# Full-size structure of BASELINE.json's configs as measured on the reference's own analysis ("Measured
# structure" table: devSlotSize, nsuper, isize, xsize, csize, etree height, flops).  A scalar fingerprint, but at the size the
# bench runs: any deviation of the ordering, the column counts or the relaxed amalgamation shows up in it.
SURVEY_TABLE = {
    "cfg1": dict(slot=5_700_000_000, nsuper=73_431, isize=2_014_421, xsize=104_784_416, csize=1_034_289, flops=1.923e10),
    "cfg2": dict(slot=5_700_000_000, nsuper=104_917, isize=4_497_623, xsize=1_096_586_362, csize=124_010_496, flops=5.905e12),
    "cfg4": dict(slot=5_380_000_000, nsuper=106_261, isize=4_508_272, xsize=1_760_091_310, csize=123_210_000, flops=1.188e13),
}


@pytest.mark.parametrize("cfg", sorted(SURVEY_TABLE))
def test_full_size_structure_matches_the_reference_measurements(cfg):
    want = SURVEY_TABLE[cfg]
    gen, g, kind = sf.matrices.CONFIGS[cfg]
    A, _ = build_matrix(gen, g)
    S = sf.analyze(A, kind, devslot=want["slot"])
    for k in ("nsuper", "isize", "xsize", "csize"):
        assert S.int(k) == want[k], k
    assert S.flops == pytest.approx(want["flops"], rel=5e-4)
    # re-using the ordering skips METIS and must give the identical structure (row f3)
    S2 = sf.analyze(A, kind, devslot=want["slot"], order=S.array("Order"))
    assert S2.double("t_order") == 0.0
    for k in ("Perm", "Super", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(S.array(k), S2.array(k)), k


def test_parallel_ordering_is_opt_in_and_equivalent():
    """sfb200_set_ordering_processes(p > 1): nested dissection with the top separators computed in the caller and the subgraphs
    ordered by METIS_NodeND in forked processes (csrc/parallel_order.cpp).  Not the reference's permutation -- which stays
    the default -- but a valid one of the same quality, and the factorization built on it is correct."""
    import sforacle
    import sparseframe_b200 as sf
    from sparseframe_b200 import _capi
    L = _capi.lib()
    n, r, c, x, sym = sf.matrices.generate("lap3d", 30)          # 27 000 vertices: one separator level is opened
    A = sf.compress(n, r, c, x, sym)
    assert L.sfb200_set_ordering_processes(1) == 1
    ref = sf.analyze(A, "cholesky")
    try:
        assert L.sfb200_set_ordering_processes(4) == 4
        par = sf.analyze(A, "cholesky")
    finally:
        L.sfb200_set_ordering_processes(1)
    again = sf.analyze(A, "cholesky")
    assert np.array_equal(again.array("Perm"), ref.array("Perm"))          # the default is untouched
    perm = par.array("Perm")
    assert np.array_equal(np.sort(perm), np.arange(n)) and not np.array_equal(perm, ref.array("Perm"))
    assert par.array("Super")[-1] == n and par.int("xsize") == par.array("Lsxp")[-1]
    assert 0.8 <= par.flops / ref.flops <= 1.2 and 0.85 <= par.int("xsize") / ref.int("xsize") <= 1.15
    # the structure built on that ordering factorizes correctly: interpreter with the ordering fed back through analyze_ordered
    S2 = sf.analyze(A, "cholesky", order=par.array("Order"))
    for k in ("Perm", "Super", "Lsip", "Lsi"):
        assert np.array_equal(S2.array(k), par.array(k)), k
    st = {k: par.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = par.int("csize"), par.int("nsleaf")
    lsx, bad = sforacle.factorize("cholesky", st, par.array("Lx"))
    assert bad == 0
    _, backward = sforacle.validate("cholesky", st, par.array("Lx"), lsx)
    assert backward <= 1e-13
