"""Complex matrices on the GPU: sfb200_zplan_* -- the FP64 engine on the real embedding
(csrc/complex_embed.hpp, csrc/zengine.cu) -- against the golden fixtures dumped from the compiled reference's
zherk/zgemm/zpotrf/ztrsm and zgemm/zgetrf_nopiv/ztrsm paths (CH.c:2060-2070, 2134-2145; LU.c:2569-2581, 2652-2663),
against the compiled reference itself at sizes that reach the blocked large-front kernels, and through the drop-in's
SparseFrame_factorize.  Collected after every other GPU test (file name)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sfref
import sparseframe_b200 as sf
from test_complex import ZGOLDEN, load, matrix_of, defined_mask, dense_permuted

pytestmark = pytest.mark.gpu
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sparse-matrix-factorization-library_b200")


@pytest.fixture(params=ZGOLDEN, ids=[os.path.basename(p)[:-4] for p in ZGOLDEN])
def zgolden(request):
    return load(request.param)


def values_of(S, gd_kind, symmetric):
    return S.array("Lx"), (S.array("Ux") if (gd_kind == "lu" and not symmetric) else None)


def test_zplan_reproduces_golden(zgolden):
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    lz, uz = values_of(S, zgolden["kind"], zgolden["symmetric"])
    ref, mask = zgolden["Lsx"], defined_mask(zgolden)
    plan = sf.ZPlan(S)
    assert plan.int("xsize") == zgolden["xsize"] and plan.int("n") == zgolden["n"]
    for _ in range(3):                      # plain issue, graph capture, graph replay of the embedded schedule
        out = plan.factorize(lz, uz)
        assert sf.entrywise_rel_diff(out[mask], ref[mask]) <= 1e-10
        assert np.abs(out[mask] - ref[mask]).max() <= 1e-12 * np.abs(ref).max()
        assert np.array_equal(out[mask] == 0, ref[mask] == 0)
        if zgolden["kind"] == "cholesky":
            assert not out[~mask].any()     # zeros above the diagonal of every panel, as loadA's memset leaves them
    again = plan.download()
    assert np.array_equal(again, out)
    plan.close()


def test_row_phase_on_the_device(zgolden):
    """LU: with and without the row rotation the factors handed back are those of the original matrix; on the purely
    imaginary diagonal the unrotated elimination meets a zero pivot and says so."""
    if zgolden["kind"] != "lu":
        pytest.skip("LU only")
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, "lu", devslot=zgolden["devslot"])
    lz, uz = values_of(S, "lu", zgolden["symmetric"])
    ref = zgolden["Lsx"]
    on = sf.ZPlan(S, row_phase=True)
    assert on.int("row_phase") == 1
    out = on.factorize(lz, uz)
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    on.close()
    off = sf.ZPlan(S, row_phase=False)
    assert off.int("row_phase") == 0
    if zgolden["phase"] >= 0.0:
        # Re a_kk is 0 up to the rounding of exp(i pi/2): a zero pivot is reported, a tiny one wrecks the factor
        try:
            out = off.factorize(lz, uz)
            assert not np.isfinite(out).all() or np.abs(out - ref).max() > 1e-6 * np.abs(ref).max()
        except RuntimeError:
            pass
    else:
        out = off.factorize(lz, uz)
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    off.close()


@pytest.mark.parametrize("nrhs", [1, 3])
def test_zsolve_and_zresidual(zgolden, nrhs):
    """The reference's complex solve / validate are `// TODO` (CH.c:3088-3131): the embedded substitution is checked against
    the dense permuted matrix instead."""
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    lz, uz = values_of(S, zgolden["kind"], zgolden["symmetric"])
    n = zgolden["n"]
    M = dense_permuted(S, zgolden["kind"], zgolden["kind"] == "cholesky")
    b = (1 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n))
    B = b if nrhs == 1 else np.stack([b * (1 + k) * np.exp(0.3j * k) for k in range(nrhs)], axis=1)
    plan = sf.ZPlan(S)
    plan.factorize(lz, uz, download=False)
    X = plan.solve(B)
    scale = np.abs(M).sum(axis=0).max() * np.abs(X).max() + np.abs(B).max()
    assert np.abs(M @ X - B).max() <= 1e-12 * scale
    if nrhs == 1:
        res = plan.residual(X, B)
        assert 0.0 <= res["residual"] <= 1e-13
    plan.close()


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind", [("lap3d", 24, "cholesky"), ("el3d", 8, "cholesky"), ("cd3d", 20, "lu"), ("lap2d", 120, "lu")])
def test_zplan_against_compiled_reference_on_large_fronts(gen, g, kind):
    """Sizes whose top fronts go through the blocked panel factorization and the 128x128 update kernels of the embedded
    (twice as wide) structure."""
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    A = sf.compress(n, r, c, z, sym)
    R = sfref.Reference(n, r, c, z, sym, kind).analyze().factorize()
    ref, st = R.array("Lsx"), R.structure()
    S = sf.analyze(A, kind)
    for k in ("Perm", "Super", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(S.array(k)[:len(st[k])], st[k]), k
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], R.long("xsize"))
    plan = sf.ZPlan(S)
    assert plan.real.int("nlarge") > 0
    lz, uz = values_of(S, kind, sym)
    out = plan.factorize(lz, uz)
    assert sf.entrywise_rel_diff(out[mask], ref[mask]) <= 1e-10
    assert np.array_equal(out[mask] == 0, ref[mask] == 0)
    b = (1 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n))
    x = plan.solve(b)
    assert plan.residual(x, b)["residual"] <= 1e-13
    print("\n%s %d %s complex: n %d, embedded factor %.1f MB, %d launches, %.2f ms" %
          (gen, g, kind, n, 8e-6 * plan.real.int("xsize_padded"), plan.real.last_launches(), plan.last_ms()))
    plan.close()
    R.close()


def test_factorize_entry_point_with_a_complex_matrix_info(zgolden):
    """SparseFrame_factorize of the drop-in library on a matrix_info whose isComplex is set and whose structure came from
    somebody else's analysis: Lsx comes back as Complex panels equal to the reference's."""
    from sparseframe_b200 import dropin
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    lz, uz = values_of(S, zgolden["kind"], zgolden["symmetric"])
    lib = dropin.Library(zgolden["kind"])
    keep = []
    lsx = np.zeros(zgolden["xsize"], dtype=np.complex128)
    m = lib.matrix_from_structure(S, lz, uz, lsx.ctypes.data, keep)
    common = dropin.common_info_struct()
    common.numGPU = 1
    assert lib.lib.SparseFrame_factorize(C.byref(common), None, C.byref(m)) == 0
    ref, mask = zgolden["Lsx"], defined_mask(zgolden)
    assert sf.entrywise_rel_diff(lsx[mask], ref[mask]) <= 1e-10
    assert lib.lib.SparseFrame_validate(C.byref(m)) == 0         # device solve: the factor is still resident
    assert 0.0 <= m.residual <= 1e-13
    lib.lib.SparseFrame_initialize_matrix(C.byref(m))            # drops the plan cached for this matrix_info


@pytest.mark.parametrize("tree,gen,g", [("Cholesky", "lap3d", 14), ("LU", "cd3d", 12)])
def test_demo_runs_on_a_complex_file(tmp_path, tree, gen, g):
    """The reference's driver, Demo/demo <file>, on a complex MatrixMarket file: prints a residual where the reference prints
    NaN (its complex solve is a TODO)."""
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(tree == "Cholesky"))
    mtx = tmp_path / "z.mtx"
    sf.matrices.write_matrix_market(str(mtx), n, r, c, z, sym)
    p = subprocess.run([os.path.join(PKG, tree, "Demo", "demo"), str(mtx)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "is complex" in p.stdout
    res = [float(line.split(":")[1]) for line in p.stdout.splitlines() if line.startswith("residual")]
    assert len(res) == 1 and 0.0 <= res[0] <= 1e-13, p.stdout


def test_zgroup_on_several_gpus(zgolden):
    """A complex matrix on two GPUs (sfb200_zgroup_*: the etree partition of the embedded structure as an in-process group,
    values expanded and complex panels read back on the host).  Needs a GPU per rank."""
    if sf.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (ranks of one process need a GPU each, DESIGN.md section 9)")
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    lz, uz = values_of(S, zgolden["kind"], zgolden["symmetric"])
    ref, mask = zgolden["Lsx"], defined_mask(zgolden)
    grp = sf.ZGroup(S, [0, 1])
    for _ in range(3):
        out = grp.factorize(lz, uz)
        assert sf.entrywise_rel_diff(out[mask], ref[mask]) <= 1e-10
        assert np.array_equal(out[mask] == 0, ref[mask] == 0)
    grp.close()
