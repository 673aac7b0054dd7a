"""Complex matrices (matrix_info->isComplex), CPU tier.

The reference factorizes them with zherk/zgemm/zpotrf/ztrsm (CH.c:2060-2070, 2134-2145) and zgemm/magma_zgetrf_nopiv/ztrsm
(LU.c:2569-2581, 2652-2663); its complex solve / validate are `// TODO` (CH.c:3088-3131).  The engine runs them as the real
embedding of csrc/complex_embed.hpp.  Checked here without a GPU:
  * the analysis with the complex element size is bit-identical to the compiled reference's and to the golden fixtures;
  * the embedded schedule on the CPU interpreter + the panel read-back reproduce the reference's complex factor;
  * the embedded substitution is the complex solve; the row-phase rotation rescues matrices whose pivots have small real parts.
"""
import glob
import os

import numpy as np
import pytest

import sforacle
import sfref
import sparseframe_b200 as sf
from conftest import ROOT

ZGOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "complex", "*.npz")))
INT_ARRAYS = ["Perm", "Parent", "Post", "ColCount", "Lp", "Li", "LTp", "LTi", "Super", "SuperMap", "Sparent", "Lsip", "Lsxp",
              "Lsi", "ST_Map", "ST_Pointer", "ST_Index", "Aoffset", "Moffset"]


def load(path):
    d = dict(np.load(path))
    gen, fk, symtxt = [str(v) for v in d["meta"]]
    n, ns, isize, xsize, csize, nstage, nsleaf, slot, g = [int(v) for v in d["scalars"]]
    d.update(gen=gen, kind=fk, symmetric=(symtxt == "symmetric"), n=n, nsuper=ns, isize=isize, xsize=xsize, csize=csize,
             nstage=nstage, nsleaf=nsleaf, devslot=slot, grid=g, phase=float(d["diag_phase"][0]))
    return d


@pytest.fixture(params=ZGOLDEN, ids=[os.path.basename(p)[:-4] for p in ZGOLDEN])
def zgolden(request):
    return load(request.param)


def matrix_of(gd):
    n, r, c, z, sym = sf.matrices.generate_complex(gd["gen"], gd["grid"], hermitian=(gd["kind"] == "cholesky"),
                                                   diag_phase=None if gd["phase"] < 0 else gd["phase"])
    return sf.compress(n, r, c, z, sym), (n, r, c, z, sym)


def defined_mask(gd):
    return sf.lower_trapezoid_mask(gd["kind"], gd["Super"], gd["Lsip"], gd["Lsxp"], gd["xsize"])


def test_there_are_complex_goldens():
    assert len(ZGOLDEN) >= 6


def test_complex_analysis_reproduces_golden(zgolden):
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    assert S.complex
    for k in INT_ARRAYS:
        got = S.array(k)
        if k == "Sparent":
            got = got[:zgolden["nsuper"]]
        assert np.array_equal(got, zgolden[k]), k
    for k in ("isize", "xsize", "csize", "nstage", "nsleaf", "nsuper"):
        assert S.int(k) == zgolden[k], k
    assert np.array_equal(S.array("LeafQueue")[:zgolden["nsleaf"]], zgolden["LeafQueue"][:zgolden["nsleaf"]])
    assert np.array_equal(S.array("Lx"), zgolden["Lx"])           # complex values move whole, never conjugated (CH.c:1048-1060)
    if "Ux" in zgolden:
        assert np.array_equal(S.array("Ux"), zgolden["Ux"])
        for k in ("Up", "Ui", "UTp", "UTi"):
            assert np.array_equal(S.array(k), zgolden[k]), k


def test_slot_caps_count_complex_bytes():
    """The same pattern and slot size give a different structure for real and for complex values: the caps are in bytes of
    the panel's element type (CH.c:1480-1496)."""
    gd = load([p for p in ZGOLDEN if "slot" in p][0])
    A, (n, r, c, z, sym) = matrix_of(gd)
    Sz = sf.analyze(A, gd["kind"], devslot=gd["devslot"])
    Sr = sf.analyze(sf.compress(n, r, c, np.abs(z), sym), gd["kind"], devslot=gd["devslot"])
    assert Sz.int("nsuper") == gd["nsuper"]
    assert Sr.int("nsuper") != Sz.int("nsuper") or Sr.int("nstage") != Sz.int("nstage")


@pytest.mark.parametrize("row_phase", [False, True])
@pytest.mark.parametrize("small", [(-1, -1), (0, 0)], ids=["default", "all_large"])
def test_embedded_schedule_reproduces_golden(zgolden, row_phase, small):
    if zgolden["phase"] >= 0.0 and not row_phase:
        pytest.skip("purely imaginary diagonal: covered by test_row_phase_rescues_an_imaginary_diagonal")
    A, _ = matrix_of(zgolden)
    out, bad, stats, _ = sforacle.check_plan_complex(zgolden["kind"], A.n, A.Cp, A.Ci, A.Cx, A.symmetric, zgolden["devslot"],
                                                     small_a=small[0], small_b=small[1], row_phase=row_phase)
    assert bad == 0
    assert stats[0] == zgolden["nsuper"] and stats[1] == zgolden["xsize"]
    ref = zgolden["Lsx"]
    mask = defined_mask(zgolden)
    assert sf.entrywise_rel_diff(out[mask], ref[mask]) <= 1e-10
    assert np.abs(out[mask] - ref[mask]).max() <= 1e-13 * np.abs(ref).max()
    if zgolden["kind"] == "cholesky":
        # the reference's panels: zeros above the diagonal (loadA's memset), a real diagonal (zpotrf)
        assert not out[~mask].any() and not ref[~mask].any()
        for s in range(zgolden["nsuper"]):
            nc, nr = zgolden["Super"][s + 1] - zgolden["Super"][s], zgolden["Lsip"][s + 1] - zgolden["Lsip"][s]
            dg = out[zgolden["Lsxp"][s]:zgolden["Lsxp"][s] + nc * nr].reshape(nc, nr)[np.arange(nc), np.arange(nc)]
            assert not dg.imag.any() and (dg.real > 0).all()


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind,devslot", [("lap3d", 9, "cholesky", sfref.DEFAULT_DEVSLOT), ("el3d", 4, "cholesky", 200_000),
                                                ("cd3d", 9, "lu", sfref.DEFAULT_DEVSLOT), ("lap3d", 6, "lu", 50_000)])
def test_embedded_schedule_reproduces_compiled_reference(gen, g, kind, devslot):
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    A = sf.compress(n, r, c, z, sym)
    R = sfref.Reference(n, r, c, z, sym, kind, devslot=devslot, blas_threads=1).analyze().factorize()
    ref, st = R.array("Lsx"), R.structure()
    S = sf.analyze(A, kind, devslot=devslot)
    for k in ("Perm", "Super", "Lsip", "Lsxp", "Lsi"):
        assert np.array_equal(S.array(k)[:len(st[k])], st[k]), k
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], R.long("xsize"))
    for row_phase in (False, True):
        out, bad, _, _ = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, sym, devslot, nb=32, outer=64, row_phase=row_phase)
        assert bad == 0
        assert np.abs(out[mask] - ref[mask]).max() <= 1e-13 * np.abs(ref).max()


def dense_permuted(S, kind, hermitian):
    """P A P^T as the factorization reads it: Cholesky -- the permuted lower triangle completed as a HERMITIAN matrix
    (zherk / zpotrf read nothing else, the diagonal's imaginary part included); LU -- lower + upper values, or A = A^T."""
    n = S.int("n")
    Lp, Li, Lx = S.array("Lp"), S.array("Li"), S.array("Lx")
    M = np.zeros((n, n), complex)
    cols = np.repeat(np.arange(n), np.diff(Lp))
    if kind == "cholesky":
        M[Li, cols] = Lx
        M = np.tril(M, -1) + np.tril(M, -1).conj().T + np.diag(np.diag(M).real)
    elif S.int("symmetric"):
        M[Li, cols] = Lx
        M = M + M.T - np.diag(np.diag(M))
    else:
        M[Li, cols] = Lx
        Up, Ui, Ux = S.array("Up"), S.array("Ui"), S.array("Ux")
        M[np.repeat(np.arange(n), np.diff(Up)), Ui] = Ux
    return M


@pytest.mark.parametrize("gen,g,kind", [("lap3d", 7, "cholesky"), ("cd3d", 7, "lu"), ("lap2d", 12, "lu")])
def test_embedded_substitution_is_the_complex_solve(gen, g, kind):
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    A = sf.compress(n, r, c, z, sym)
    S = sf.analyze(A, kind)
    M = dense_permuted(S, kind, kind == "cholesky")
    b = (1 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n))        # the reference's complex Bx (CH.c:3190-3191)
    for row_phase in (False, True):
        _, bad, _, x = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, sym, sf.DEFAULT_DEVSLOT, row_phase=row_phase, rhs=b)
        assert bad == 0
        assert np.abs(M @ x - b).max() <= 1e-13 * (np.abs(M).sum(axis=0).max() * np.abs(x).max() + np.abs(b).max())


def test_row_phase_rescues_an_imaginary_diagonal(zgolden):
    """A purely IMAGINARY dominant diagonal: the pivots of the embedded elimination are the real parts of the complex pivots,
    here 0 at the first column -- without the rotation the elimination stops (or the factor is garbage); with it the result is
    the reference's."""
    if zgolden["phase"] < 0.0:
        pytest.skip("only the imaginary-diagonal fixture")
    A, _ = matrix_of(zgolden)
    ref = zgolden["Lsx"]
    good, bad, _, _ = sforacle.check_plan_complex("lu", A.n, A.Cp, A.Ci, A.Cx, A.symmetric, zgolden["devslot"], row_phase=True)
    assert bad == 0 and np.abs(good - ref).max() <= 1e-13 * np.abs(ref).max()
    plain, bad0, _, _ = sforacle.check_plan_complex("lu", A.n, A.Cp, A.Ci, A.Cx, A.symmetric, zgolden["devslot"], row_phase=False)
    assert bad0 > 0 or not np.isfinite(plain).all() or np.abs(plain - ref).max() > 1e-6 * np.abs(ref).max()


# ---- the drop-in library on complex input (no device needed: reader, analysis, and the host loops of solve / validate) ------

def test_dropin_completes_the_reference_todo_solve_and_validate(zgolden):
    """SparseFrame_validate on a complex factor: the reference leaves the substitution and the norms as `// TODO`
    (CH.c:3088-3131, 3229-3256) and prints NaN; the drop-in carries them out.  Fed with the REFERENCE's own factor (golden
    Lsx), no device involved: the residual of the permuted system must be at rounding level."""
    import ctypes as C
    from sparseframe_b200 import dropin
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    lib = dropin.Library(zgolden["kind"])
    keep = []
    lsx = np.ascontiguousarray(zgolden["Lsx"])
    ux = S.array("Ux") if (zgolden["kind"] == "lu" and not zgolden["symmetric"]) else None
    m = lib.matrix_from_structure(S, S.array("Lx"), ux, lsx.ctypes.data, keep)
    assert m.isComplex == 1
    assert lib.lib.SparseFrame_validate(C.byref(m)) == 0
    assert 0.0 <= m.residual <= 1e-14
    n = zgolden["n"]
    x = np.ctypeslib.as_array(C.cast(m.Xx, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
    b = np.ctypeslib.as_array(C.cast(m.Bx, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
    assert np.allclose(b, (1 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n)))      # CH.c:3190-3191
    M = dense_permuted(S, zgolden["kind"], zgolden["kind"] == "cholesky")
    assert np.abs(M @ x - b).max() <= 1e-13 * (np.abs(M).sum(axis=0).max() * np.abs(x).max() + 1)


def test_dropin_reads_and_analyzes_a_complex_file(zgolden, tmp_path):
    """SparseFrame_read_matrix + SparseFrame_analyze of the drop-in on a complex MatrixMarket file leave the reference's
    matrix_info: isComplex set, structure and permuted complex values bit-identical to the golden fixture."""
    import subprocess
    from test_reader import PKG
    probe = r"""
#include <stdio.h>
#include <string.h>
#include "SparseFrame.h"
int main(int argc, char **argv) {
  static struct matrix_info_struct m;
  static struct common_info_struct c;
  memset(&m, 0, sizeof(m)); memset(&c, 0, sizeof(c));
  c.devSlotSize = strtoul(argv[3], 0, 10);
  m.path = argv[1];
  SparseFrame_initialize_matrix(&m);
  if (SparseFrame_read_matrix(&m)) return 3;
  if (SparseFrame_analyze(&c, &m)) return 4;
  FILE *f = fopen(argv[2], "wb");
  long h[6] = {m.nrow, m.nsuper, m.isize, m.nzmax, m.isComplex, m.xsize};
  fwrite(h, sizeof(long), 6, f);
  fwrite(m.Perm, sizeof(long), m.nrow, f);
  fwrite(m.Super, sizeof(long), m.nsuper + 1, f);
  fwrite(m.Lsip, sizeof(long), m.nsuper + 1, f);
  fwrite(m.Lsxp, sizeof(long), m.nsuper + 1, f);
  fwrite(m.Lsi, sizeof(long), m.isize, f);
  fwrite(m.Lx, sizeof(Complex), m.Lp[m.nrow], f);
  fclose(f);
  return 0; }
"""
    tree = "LU" if zgolden["kind"] == "lu" else "Cholesky"
    src, exe = tmp_path / "probe.c", tmp_path / "probe"
    src.write_text("#include <stdlib.h>\n" + probe)
    lib = os.path.join(PKG, tree, "Lib")
    subprocess.check_call(["gcc", "-fopenmp", "-I" + os.path.join(PKG, tree, "Include"), str(src), "-o", str(exe),
                           "-L" + lib, "-lSparseFrame", "-Wl,-rpath," + lib])
    _, (n, r, c, z, sym) = matrix_of(zgolden)
    mtx, out = tmp_path / "z.mtx", tmp_path / "z.dump"
    sf.matrices.write_matrix_market(str(mtx), n, r, c, z, sym)
    subprocess.run([str(exe), str(mtx), str(out), str(zgolden["devslot"])], check=True, capture_output=True)
    raw = np.fromfile(out, dtype=np.int64)
    nrow, ns, isize, nz, cplx, xsize = raw[:6]
    assert (nrow, ns, isize, cplx, xsize) == (n, zgolden["nsuper"], zgolden["isize"], 1, zgolden["xsize"])
    at = 6
    for name, cnt in (("Perm", n), ("Super", ns + 1), ("Lsip", ns + 1), ("Lsxp", ns + 1), ("Lsi", isize)):
        assert np.array_equal(raw[at:at + cnt], zgolden[name][:cnt]), name
        at += cnt
    assert np.array_equal(raw[at:].view(np.complex128), zgolden["Lx"])


# ---- irregular complex inputs: random sparsity in shuffled file order, duplicates, forests, tiny matrices -------------------

def complexify(r, c, x, hermitian, seed):
    rng = np.random.default_rng(seed)
    z = x * np.exp(1j * rng.uniform(-0.6, 0.6, len(x)))
    if hermitian:
        z[r == c] = x[r == c]
    return z


def check_complex_against_reference(n, r, c, z, sym, kind, slot=sf.DEFAULT_DEVSLOT, **prm):
    A = sf.compress(n, r, c, z, sym)
    R = sfref.Reference(n, r, c, z, sym, kind=kind, devslot=slot, blas_threads=1).analyze().factorize()
    S = sf.analyze(A, kind, devslot=slot)
    for k in ("Perm", "Parent", "Post", "ColCount", "Lp", "Li", "Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "ST_Map"):
        assert np.array_equal(S.array(k), R.array(k)[:len(S.array(k))]), k
    assert np.array_equal(S.array("Lx"), R.array("Lx"))
    ref = R.array("Lsx")
    mask = sf.lower_trapezoid_mask(kind, R.array("Super"), R.array("Lsip"), R.array("Lsxp"), R.long("xsize"))
    for row_phase in ((False, True) if kind == "lu" else (False,)):
        out, bad, _, _ = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, sym, slot, row_phase=row_phase, **prm)
        assert bad == 0
        assert sf.entrywise_rel_diff(out[mask], ref[mask]) <= 1e-10
        assert np.array_equal(out[mask] == 0, ref[mask] == 0)
    R.close()


@pytest.mark.skipif(not sfref.available("cholesky"), reason="compiled reference not built")
@pytest.mark.parametrize("n,density,seed", [(40, 0.1, 11), (250, 0.02, 12), (700, 0.006, 13)])
@pytest.mark.parametrize("prm", [dict(), dict(nb=8, outer=16, small_a=0, small_b=0)], ids=["default", "nb8"])
def test_random_hermitian_cholesky(n, density, seed, prm):
    from test_irregular import random_matrix
    n, r, c, x, sym = random_matrix(n, density, seed, True)
    check_complex_against_reference(n, r, c, complexify(r, c, x, True, seed), sym, "cholesky", **prm)


@pytest.mark.skipif(not sfref.available("lu"), reason="compiled reference not built")
@pytest.mark.parametrize("n,density,seed,patsym", [(40, 0.1, 14, True), (250, 0.02, 15, False), (600, 0.006, 16, False)])
@pytest.mark.parametrize("prm", [dict(), dict(nb=8, outer=16, small_a=0, small_b=0)], ids=["default", "nb8"])
def test_random_complex_lu(n, density, seed, patsym, prm):
    from test_irregular import random_matrix
    n, r, c, x, sym = random_matrix(n, density, seed, False, patsym)
    check_complex_against_reference(n, r, c, complexify(r, c, x, False, seed), sym, "lu", **prm)


@pytest.mark.skipif(not sfref.available("lu"), reason="compiled reference not built")
@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_complex_duplicates_last_one_wins(kind, monkeypatch):
    """Duplicate (i, j) entries: nothing is summed, loadA keeps the last copy (CH.c:2018, LU.c:2498/2513) -- of the diagonal
    too, whose phase the row rotation is taken from.  (The switch reproduces the reference's METIS input, repeated neighbours
    included, so that the structures can be compared; tests/test_plan.py::test_heavily_duplicated_entries_do_not_reach_metis.)"""
    monkeypatch.setenv("SFB200_METIS_KEEP_DUPLICATES", "1")
    from test_irregular import random_matrix
    n, r, c, x, sym = random_matrix(120, 0.04, 21, kind == "cholesky")
    z = complexify(r, c, x, kind == "cholesky", 21)
    rng = np.random.default_rng(22)
    pick = rng.choice(len(r), 40, replace=False)
    # earlier copies with other values in front of the real ones (diagonal entries among them)
    r2 = np.concatenate([r[pick], r])
    c2 = np.concatenate([c[pick], c])
    junk = 0.3 * z[pick] * np.exp(0.7j) if kind == "lu" else 0.3 * z[pick]
    z2 = np.concatenate([junk, z])
    check_complex_against_reference(n, r2, c2, z2, sym, kind)


@pytest.mark.skipif(not sfref.available("lu"), reason="compiled reference not built")
@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_complex_forest_and_tiny(kind):
    parts, off = [], 0
    for gen, g in (("lap2d", 7), ("lap3d", 4), ("lap2d", 3)):
        n, r, c, z, _ = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
        parts.append((r + off, c + off, z))
        off += n
    iso = np.arange(off, off + 4)
    parts.append((iso, iso, np.full(4, 2.5 + 0j) if kind == "cholesky" else np.full(4, 2.5 * np.exp(0.4j))))
    n = off + 4
    r, c, z = (np.concatenate([p[k] for p in parts]) for k in range(3))
    check_complex_against_reference(n, r, c, z, True, kind)
    # 1 x 1 and a dense 2 x 2 (the reference's own analysis of a two-vertex graph does not return a permutation, so these are
    # checked against numpy instead: one supernode, identity or swap ordering)
    for n1, vals in ((1, [3.0 + 0j]), (2, [2.0 + 0j, 0.5 - 0.25j, 3.0 + 0j])):
        rr = np.array([0] if n1 == 1 else [0, 1, 1])
        cc = np.array([0] if n1 == 1 else [0, 0, 1])
        zz = np.array(vals) if kind == "cholesky" else np.array(vals) * np.exp(0.3j)
        A = sf.compress(n1, rr, cc, zz, True)
        S = sf.analyze(A, kind)
        M = dense_permuted(S, kind, kind == "cholesky")
        out, bad, _, x = sforacle.check_plan_complex(kind, n1, A.Cp, A.Ci, A.Cx, True, sf.DEFAULT_DEVSLOT,
                                                     rhs=np.arange(1, n1 + 1) * (1 + 0.5j))
        assert bad == 0
        assert np.abs(M @ x - np.arange(1, n1 + 1) * (1 + 0.5j)).max() <= 1e-14
        if kind == "cholesky" and S.int("nsuper") == 1:
            Lc = np.linalg.cholesky(M)
            got = out.reshape(n1, n1).T
            assert np.abs(np.tril(got) - Lc).max() <= 1e-14


def test_zplan_needs_a_device_and_a_complex_structure():
    """No CPU fallback: without a CUDA device the numeric entry point fails loudly; a real structure is refused."""
    import ctypes as C
    from sparseframe_b200 import _capi
    L = _capi.lib()
    n, r, c, z, sym = sf.matrices.generate_complex("lap2d", 6, hermitian=True)
    Sz = sf.analyze(sf.compress(n, r, c, z, sym), "cholesky")
    Sr = sf.analyze(sf.compress(n, r, c, np.abs(z), sym), "cholesky")
    h = C.c_void_p()
    assert L.sfb200_zplan_create(Sr._h, 0, C.byref(h)) == 1 and b"complex" in L.sfb200_last_error()      # SFB200_ERR_ARG
    if sf.device_count() == 0:
        assert L.sfb200_zplan_create(Sz._h, 0, C.byref(h)) == 3 and b"no CPU fallback" in L.sfb200_last_error()
        with pytest.raises(RuntimeError):
            sf.ZPlan(Sz)


# ---- the plain-C restatement of the reference's isComplex branches (oracle/sf_oracle_z.c) is pinned here ------------------

def structure_of(gd):
    st = {k: gd[k] for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = gd["csize"], gd["nsleaf"]
    if "Up" in gd:
        st["Up"], st["Ui"] = gd["Up"], gd["Ui"]
    return st


def test_complex_restatement_reproduces_golden(zgolden):
    st = structure_of(zgolden)
    out, bad = sforacle.zfactorize(zgolden["kind"], st, zgolden["Lx"], zgolden.get("Ux"))
    assert bad == 0
    ref, mask = zgolden["Lsx"], defined_mask(zgolden)
    assert np.abs(out[mask] - ref[mask]).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(out == 0, ref == 0)                    # the panels' zero pattern, undefined corner included
    res, norms, x = sforacle.zvalidate(zgolden["kind"], st, zgolden["Lx"], out, zgolden.get("Ux"))
    assert 0.0 <= res <= 1e-14
    # ... and the solve it completes agrees with the dense permuted matrix
    A, _ = matrix_of(zgolden)
    S = sf.analyze(A, zgolden["kind"], devslot=zgolden["devslot"])
    M = dense_permuted(S, zgolden["kind"], zgolden["kind"] == "cholesky")
    n = zgolden["n"]
    b = (1 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n))
    assert np.abs(M @ x - b).max() <= 1e-13 * (np.abs(M).sum(axis=0).max() * np.abs(x).max() + 1)
    assert abs(norms[0] - np.abs(M).sum(axis=0).max()) <= 1e-12 * norms[0]


@pytest.mark.skipif(not sfref.available("lu"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind", [("lap3d", 8, "cholesky"), ("el3d", 3, "cholesky"), ("cd3d", 8, "lu"), ("lap2d", 15, "lu")])
def test_complex_restatement_reproduces_compiled_reference(gen, g, kind):
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    R = sfref.Reference(n, r, c, z, sym, kind, blas_threads=1).analyze()
    leaf = R.array("LeafQueue")
    R.factorize()
    st = {k: R.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "Lp", "Li")}
    st["LeafQueue"], st["csize"], st["nsleaf"] = leaf, R.long("csize"), R.long("nsleaf")
    uz = None
    if kind == "lu" and not sym:
        st["Up"], st["Ui"], uz = R.array("Up"), R.array("Ui"), R.array("Ux")
    out, bad = sforacle.zfactorize(kind, st, R.array("Lx"), uz)
    ref = R.array("Lsx")
    assert bad == 0 and np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(out == 0, ref == 0)
    R.close()


# ---- complex matrices on several ranks (sfb200_zgroup_*: the partition of the embedded structure + host expand / read-back) ---

@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_embedding_reproduces_golden(zgolden, world, monkeypatch):
    """The composition sfb200_zgroup_factorize runs -- expanded values, the etree partition of the EMBEDDED structure (group-
    factorized panels and split Schur updates forced by small thresholds), complex panels read out of the packed embedded
    factor -- on the CPU: lock step in both rank orders against the reference's golden factor."""
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "32")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "16")
    A, _ = matrix_of(zgolden)
    ref, mask = zgolden["Lsx"], defined_mask(zgolden)
    for reverse in (False, True):
        if reverse:
            monkeypatch.setenv("SFCHECK_RANKS_REVERSED", "1")
        out, bad = sforacle.check_plan_complex_partitioned(zgolden["kind"], A.n, A.Cp, A.Ci, A.Cx, A.symmetric, zgolden["devslot"],
                                                           world, nb=8, outer=16, small_a=0, small_b=0, row_phase=True,
                                                           xsize=zgolden["xsize"])
        assert bad == 0
        assert np.abs(out[mask] - ref[mask]).max() <= 1e-13 * np.abs(ref).max()
        assert np.array_equal(out[mask] == 0, ref[mask] == 0)


@pytest.mark.skipif(not sfref.available("lu"), reason="compiled reference not built")
@pytest.mark.parametrize("gen,g,kind,world", [("lap3d", 11, "cholesky", 4), ("lap3d", 12, "cholesky", 8), ("cd3d", 10, "lu", 5),
                                              ("lap2d", 40, "lu", 8)])
def test_partitioned_embedding_reproduces_compiled_reference(gen, g, kind, world, monkeypatch):
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "64")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "32")
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    A = sf.compress(n, r, c, z, sym)
    R = sfref.Reference(n, r, c, z, sym, kind, blas_threads=1).analyze().factorize()
    ref = R.array("Lsx")
    mask = sf.lower_trapezoid_mask(kind, R.array("Super"), R.array("Lsip"), R.array("Lsxp"), R.long("xsize"))
    for row_phase in ((False, True) if kind == "lu" else (False,)):
        out, bad = sforacle.check_plan_complex_partitioned(kind, n, A.Cp, A.Ci, A.Cx, sym, sf.DEFAULT_DEVSLOT, world,
                                                           row_phase=row_phase, xsize=R.long("xsize"))
        assert bad == 0
        assert np.abs(out[mask] - ref[mask]).max() <= 1e-13 * np.abs(ref).max()
    R.close()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_partitioned_embedding_under_the_threaded_protocol(zgolden, world, monkeypatch):
    """The exchange protocol of the partition sfb200_zgroup builds (plans of the embedded structure): one free-running thread
    per rank, the plan's own flag sequence numbers, two factorizations back to back, with and without jitter -- on the small
    fixtures some ranks own almost nothing, the case a lock-step simulation cannot hang on."""
    import ctypes as C
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "32")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "16")
    f = sforacle._load_chk().sfcheck_factorize_complex_threaded
    f.restype = C.c_long
    f.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + [C.c_int] * 7 + \
        [C.c_double, C.c_void_p]
    A, _ = matrix_of(zgolden)
    if zgolden["phase"] >= 0.0:
        pytest.skip("the purely imaginary diagonal needs the row rotation (covered by the lock-step test)")
    ref, mask = zgolden["Lsx"], defined_mask(zgolden)
    Cz = np.ascontiguousarray(A.Cx, dtype=np.complex128)
    for jitter in (0, 25):
        out = np.zeros(zgolden["xsize"], dtype=np.complex128)
        rc = f(int(zgolden["kind"] == "lu"), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data, Cz.ctypes.data,
               zgolden["devslot"], 8, 16, 0, 0, world, 2, jitter, 60.0, out.ctypes.data)
        assert rc == 0
        assert np.abs(out[mask] - ref[mask]).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("kind", ["cholesky", "lu"])
def test_empty_and_diagonal_complex_matrices(kind):
    """n = 0 and a matrix without off-diagonal entries (METIS skipped, every column its own chain, CH.c:933-939)."""
    e = np.zeros(0, dtype=np.int64)
    A = sf.compress(0, e, e, np.zeros(0, dtype=complex), True)
    S = sf.analyze(A, kind)
    assert S.complex and S.int("nsuper") == 0 and S.int("xsize") == 0
    out, bad, stats, _ = sforacle.check_plan_complex(kind, 0, A.Cp, A.Ci, A.Cx, True, sf.DEFAULT_DEVSLOT)
    assert out.size == 0 and bad == 0
    n = 7
    d = np.arange(1.0, n + 1) * (np.exp(0.3j) if kind == "lu" else 1.0)
    A = sf.compress(n, np.arange(n), np.arange(n), d.astype(complex), kind == "cholesky")
    S = sf.analyze(A, kind)
    out, bad, _, x = sforacle.check_plan_complex(kind, n, A.Cp, A.Ci, A.Cx, kind == "cholesky", sf.DEFAULT_DEVSLOT,
                                                 row_phase=True, rhs=np.ones(n, complex))
    assert bad == 0
    dp = d[S.array("Perm")]
    assert np.abs(x * dp - 1).max() <= 1e-15
    assert np.abs(out - (np.sqrt(dp.real) if kind == "cholesky" else dp)).max() <= 1e-15


@pytest.mark.parametrize("gen,g,kind,world,prm", [
    ("lap3d", 16, "cholesky", 1, dict(nb=8, outer=32 * 100000 + 16, small_a=0, small_b=0)),
    ("cd3d", 12, "lu", 1, dict(nb=8, outer=32 * 100000 + 16, small_a=0, small_b=0)),
    ("lap3d", 24, "cholesky", 1, dict()),
    ("lap3d", 14, "cholesky", 4, dict(nb=8, outer=16, small_a=0, small_b=0)),
    ("cd3d", 10, "lu", 3, dict(nb=8, outer=16, small_a=0, small_b=0))])
def test_embedded_schedules_are_race_free(gen, g, kind, world, prm, monkeypatch):
    """The happens-before check of tests/test_race_check.py on the schedules a complex matrix is replayed with: the embedded
    plan with both look-aheads (three blocking levels), and every rank's plan of a partition."""
    import ctypes as C
    monkeypatch.setenv("SFCHECK_DIST_PANEL_MIN", "64")
    monkeypatch.setenv("SFCHECK_DIST_MIN_M", "32")
    f = sforacle._load_chk().sfcheck_race_check_complex
    f.restype = C.c_long
    f.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + [C.c_int] * 5
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=(kind == "cholesky"))
    A = sf.compress(n, r, c, z, sym)
    hazards = f(int(kind == "lu"), int(sym), n, A.Cp.ctypes.data, A.Ci.ctypes.data, A.Cx.ctypes.data, sf.DEFAULT_DEVSLOT,
                prm.get("nb", 0), prm.get("outer", 0), prm.get("small_a", -1), prm.get("small_b", -1), world)
    assert hazards == 0
