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
