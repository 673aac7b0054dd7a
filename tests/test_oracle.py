"""The oracle (oracle/sf_oracle.c, plain-C restatement of the reference's numeric path) pinned against
(a) the committed golden vectors dumped from the compiled reference and (b) the compiled reference itself."""
import numpy as np
import pytest

import sforacle
import sfref
import sparseframe_b200 as sf

TOL = 1e-11   # restatement vs reference: same algorithm, textbook BLAS vs OpenBLAS summation order


def _structure(g):
    st = {k: g[k] for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = g["csize"], g["nsleaf"]
    if "Up" in g:
        st["Up"], st["Ui"] = g["Up"], g["Ui"]
    return st


def test_oracle_matches_golden_factor(golden):
    st = _structure(golden)
    lsx, bad = sforacle.factorize(golden["kind"], st, golden["Lx"], golden.get("Ux"))
    assert bad == 0
    mask = sf.lower_trapezoid_mask(golden["kind"], golden["Super"], golden["Lsip"], golden["Lsxp"], golden["xsize"])
    a, b = golden["Lsx"][mask], lsx[mask]
    assert sf.entrywise_rel_diff(a, b) <= TOL
    assert np.array_equal(a == 0, b == 0), "explicit zeros of relaxed amalgamation must stay exactly zero"


def test_oracle_residual_matches_golden(golden):
    st = _structure(golden)
    res, backward = sforacle.validate(golden["kind"], st, golden["Lx"], golden["Lsx"], golden.get("Ux"))
    # same factor, same formula (CH.c:3262) -> same number up to the order of the residual sums
    assert abs(res - float(golden["residual"][0])) <= 1e-15
    assert backward <= 1e-12


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind", [("lap2d", 40, "cholesky"), ("lap3d", 14, "cholesky"), ("el3d", 5, "cholesky"),
                                        ("cd3d", 12, "lu"), ("lap3d", 10, "lu")])
def test_oracle_matches_compiled_reference(gen, g, kind):
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    ref = sfref.Reference(n, r, c, x, sym, kind=kind, blas_threads=1).analyze().factorize()
    st = {k: ref.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = ref.long("csize"), ref.long("nsleaf")
    ux = None
    if kind == "lu" and not sym:
        st["Up"], st["Ui"] = ref.array("Up"), ref.array("Ui")
        ux = ref.array("Ux")
    lsx, bad = sforacle.factorize(kind, st, ref.array("Lx"), ux)
    assert bad == 0
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], ref.long("xsize"))
    a, b = ref.array("Lsx")[mask], lsx[mask]
    assert sf.entrywise_rel_diff(a, b) <= TOL
    assert np.array_equal(a == 0, b == 0)
    # solve restatement against the reference's own solve
    res_ref = ref.validate()
    x_ref = ref.array("Xx")
    b_vec = 1.0 + np.arange(n) / n
    x_or = sforacle.solve(kind, st, ref.array("Lsx"), b_vec)
    assert np.max(np.abs(x_or - x_ref)) <= 1e-12 * np.max(np.abs(x_ref))
    res_or, _ = sforacle.validate(kind, st, ref.array("Lx"), ref.array("Lsx"), ux)
    assert abs(res_or - res_ref) <= 1e-15
    ref.close()


def test_oracle_flags_indefinite_matrix():
    # 2x2 [[1,2],[2,1]] is not positive definite: the restated dpotrf must report it (the reference ignores info)
    st = dict(Super=[0, 2], SuperMap=[0, 0], Lsip=[0, 2], Lsxp=[0, 4], Lsi=[0, 1], LeafQueue=[0], Lp=[0, 2, 3],
              Li=[0, 1, 1], csize=0, nsleaf=1)
    _, bad = sforacle.factorize("cholesky", st, [1.0, 2.0, 1.0])
    assert bad == 1
