"""Input side of the drop-in library: SparseFrame_read_matrix on MatrixMarket text (whole-file, multi-threaded
parse) and on the binary triplet container must leave exactly the Cp/Ci/Cx that the reference's reader + compress build
(CH.c:400-587): file order inside a column, explicit zeros dropped, nothing summed, values bit-identical to "%lg"."""
import os
import subprocess

import numpy as np
import pytest

import sparseframe_b200 as sf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-matrix-factorization-library_b200")

# reads argv[1] through the drop-in's own entry points and dumps what it found
PROBE = r"""
#include <stdio.h>
#include <string.h>
#include "SparseFrame.h"
int SparseFrame_initialize_matrix(struct matrix_info_struct *);
int SparseFrame_read_matrix(struct matrix_info_struct *);
int main(int argc, char **argv) {
  static struct matrix_info_struct m;
  memset(&m, 0, sizeof(m));
  m.path = argv[1];
  SparseFrame_initialize_matrix(&m);
  int rc = SparseFrame_read_matrix(&m);
  if (rc) { printf("rc=%d\n", rc); return 3; }
  FILE *f = fopen(argv[2], "wb");
  long h[5] = {m.nrow, m.ncol, m.nzmax, m.isSymmetric, m.isComplex};
  fwrite(h, sizeof(long), 5, f);
  fwrite(m.Cp, sizeof(long), m.ncol + 1, f);
  fwrite(m.Ci, sizeof(long), m.nzmax, f);
  fwrite(m.Cx, sizeof(double), m.nzmax * (m.isComplex ? 2 : 1), f);
  fclose(f);
  return 0; }
"""


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    d = tmp_path_factory.mktemp("reader")
    src = d / "probe.c"
    src.write_text(PROBE)
    exe = d / "probe"
    lib = os.path.join(PKG, "LU", "Lib")
    subprocess.check_call(["gcc", "-fopenmp", "-I" + os.path.join(PKG, "LU", "Include"), str(src), "-o", str(exe),
                           "-L" + lib, "-lSparseFrame", "-Wl,-rpath," + lib])

    def run(path, threads=None):
        out = str(path) + ".dump"
        env = dict(os.environ)
        if threads is not None:
            env["SF_READ_THREADS"] = str(threads)
        p = subprocess.run([str(exe), str(path), out], env=env, capture_output=True, text=True)
        if p.returncode != 0:
            return None, p.stdout
        raw = np.fromfile(out, dtype=np.int64)
        nrow, ncol, nz, sym, cplx = raw[:5]
        cp = raw[5:5 + ncol + 1]
        ci = raw[5 + ncol + 1:5 + ncol + 1 + nz]
        cx = raw[5 + ncol + 1 + nz:].view(np.float64)
        return dict(nrow=nrow, ncol=ncol, nz=nz, sym=sym, cplx=cplx, Cp=cp, Ci=ci, Cx=cx), p.stdout
    return run


def expected(n, r, c, x, sym):
    A = sf.compress(n, r, c, x, sym)
    return A.Cp, A.Ci, A.Cx


@pytest.mark.parametrize("gen,g", [("lap2d", 30), ("cd3d", 9), ("el3d", 4)])
@pytest.mark.parametrize("threads", [1, 3, None])
def test_text_and_binary_give_the_reference_csc(probe, tmp_path, gen, g, threads):
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    rng = np.random.default_rng(7)
    x = x * (1.0 + rng.random(len(x)))                   # 17 significant digits: exercises the strtod path
    x[::7] = np.round(x[::7], 3)                         # short decimals: the fast path
    mtx, binf = tmp_path / "a.mtx", tmp_path / "a.sfbin"
    sf.matrices.write_matrix_market(str(mtx), n, r, c, x, sym)
    sf.matrices.write_binary(str(binf), n, r, c, x, sym)
    cp, ci, cx = expected(n, r, c, x, sym)
    for path in (mtx, binf):
        got, out = probe(path, threads)
        assert got is not None, out
        assert (got["nrow"], got["ncol"], got["sym"], got["cplx"]) == (n, n, int(sym), 0)
        assert np.array_equal(got["Cp"], cp) and np.array_equal(got["Ci"], ci)
        assert np.array_equal(got["Cx"].view(np.int64), cx.view(np.int64))          # bit-identical values


@pytest.mark.parametrize("gen,g,hermitian", [("lap3d", 6, True), ("cd3d", 5, False)])
def test_complex_text_and_binary_give_the_reference_csc(probe, tmp_path, gen, g, hermitian):
    """field "complex": (re, im) pairs, stored whole (CH.c:505-511, 572-575); an entry is dropped only if both parts are zero."""
    n, r, c, z, sym = sf.matrices.generate_complex(gen, g, hermitian=hermitian)
    z = z.copy()
    z[3] = 0.0                      # dropped
    z[5] = 1j * z[5].imag           # kept: zero real part, non-zero imaginary part
    mtx, binf = tmp_path / "z.mtx", tmp_path / "z.sfbin"
    sf.matrices.write_matrix_market(str(mtx), n, r, c, z, sym)
    sf.matrices.write_binary(str(binf), n, r, c, z, sym)
    keep = z != 0
    A = sf.compress(n, r[keep], c[keep], z[keep], sym)
    for path in (mtx, binf):
        for threads in (1, 3):
            got, out = probe(path, threads)
            assert got is not None, out
            assert (got["nrow"], got["ncol"], got["sym"], got["cplx"], got["nz"]) == (n, n, int(sym), 1, int(keep.sum()))
            assert np.array_equal(got["Cp"], A.Cp) and np.array_equal(got["Ci"], A.Ci)
            assert np.array_equal(got["Cx"].view(np.int64), A.Cx.view(np.float64).view(np.int64))


def test_reader_semantics(probe, tmp_path):
    """Comments, blank lines, explicit zeros (dropped), a fourth column, exponents, signs, duplicates kept in file order."""
    text = "\n".join([
        "%%MatrixMarket matrix coordinate real general", "% a comment", "%another", "3 3 9",
        "1 1 2.5", "2 1 -1e0", "3 1 0", "", "1 2 +0.125E+1 0", "2 2 4", "2 2 5.000000000000000001", "3 3 1e-3",
        "  3   2\t7.25  ", "1 3 0.0 1.0", "3 3 123456789012345678e-17"]) + "\n"
    p = tmp_path / "s.mtx"
    p.write_text(text)
    got, out = probe(p, 2)
    assert got is not None, out
    assert got["sym"] == 0 and got["nz"] == 9
    assert got["Cp"].tolist() == [0, 2, 6, 9]
    assert got["Ci"].tolist() == [0, 1, 0, 1, 1, 2, 2, 0, 2]
    want = [2.5, -1.0, 1.25, 4.0, float("5.000000000000000001"), 7.25, 1e-3, 0.0, float("123456789012345678e-17")]
    assert got["Cx"].tolist() == want


@pytest.mark.parametrize("body,msg", [
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1 1\n2 2 1\n", "nzmax exceeded"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1\n", "invalid matrix entry"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 3 1.0\n", "invalid matrix entry"),
    ("%%NotMatrixMarket\n", "Matrix format error"),
    ("%%MatrixMarket matrix coordinate real general\n2 2\n", "Matrix format error"),
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0\n", "imaginary part not present"),
    # every field other than "real" is complex to the reference's reader (CH.c:440-445): a pattern file has no values at all
    ("%%MatrixMarket matrix coordinate pattern general\n1 1 1\n1 1\n", "invalid matrix entry"),
    # a damaged size line: an absurd entry count is only a hint, an absurd dimension is refused instead of crashing
    ("%%MatrixMarket matrix coordinate real general\n2 2 999999999999999999\n1 1 1.0\n2 2 1.0\n", None),
    ("%%MatrixMarket matrix coordinate real general\n999999999999999 999999999999999 1\n1 1 1.0\n", "cannot allocate"),
])
def test_reader_errors(probe, tmp_path, body, msg):
    p = tmp_path / "bad.mtx"
    p.write_text(body)
    got, out = probe(p)
    if msg is None:                       # readable after all
        assert got is not None and got["nz"] == 2
        return
    assert got is None and msg in out


# analysis through the drop-in (no device needed: the plan upload is skipped with a message), dumping the structure
ANALYZE_PROBE = r"""
#include <stdio.h>
#include <string.h>
#include "SparseFrame.h"
int SparseFrame_initialize_matrix(struct matrix_info_struct *);
int SparseFrame_read_matrix(struct matrix_info_struct *);
int SparseFrame_analyze(struct common_info_struct *, struct matrix_info_struct *);
int main(int argc, char **argv) {
  static struct matrix_info_struct m;
  static struct common_info_struct c;
  memset(&m, 0, sizeof(m)); memset(&c, 0, sizeof(c));
  c.devSlotSize = 5380000000ul;
  m.path = argv[1];
  SparseFrame_initialize_matrix(&m);
  if (SparseFrame_read_matrix(&m)) return 3;
  if (SparseFrame_analyze(&c, &m)) return 4;
  FILE *f = fopen(argv[2], "wb");
  long h[4] = {m.nrow, m.nsuper, m.isize, m.nzmax};
  fwrite(h, sizeof(long), 4, f);
  fwrite(m.Perm, sizeof(long), m.nrow, f);
  fwrite(m.Post, sizeof(long), m.nrow, f);
  fwrite(m.Super, sizeof(long), m.nsuper + 1, f);
  fwrite(m.Lsip, sizeof(long), m.nsuper + 1, f);
  fwrite(m.Lsi, sizeof(long), m.isize, f);
  fwrite(m.Li, sizeof(long), m.nzmax, f);
  fwrite(m.Lx, sizeof(double), m.nzmax, f);
  fclose(f);
  return 0; }
"""


def test_ordering_cache_reproduces_the_analysis(tmp_path):
    """SF_ORDER_CACHE: the second analysis of a pattern skips METIS and must leave a bit-identical matrix_info; a damaged
    cache file is ignored."""
    src = tmp_path / "probe.c"
    src.write_text(ANALYZE_PROBE)
    exe = tmp_path / "probe"
    lib = os.path.join(PKG, "LU", "Lib")
    subprocess.check_call(["gcc", "-fopenmp", "-I" + os.path.join(PKG, "LU", "Include"), str(src), "-o", str(exe),
                           "-L" + lib, "-lSparseFrame", "-Wl,-rpath," + lib])
    n, r, c, x, sym = sf.matrices.generate("cd3d", 14)
    mtx = tmp_path / "a.mtx"
    sf.matrices.write_matrix_market(str(mtx), n, r, c, x, sym)
    cache = tmp_path / "cache"
    cache.mkdir()

    def run(tag, env_extra):
        out = tmp_path / (tag + ".dump")
        env = dict(os.environ, **env_extra)
        subprocess.run([str(exe), str(mtx), str(out)], env=env, check=True, capture_output=True)
        return out.read_bytes()

    plain = run("plain", {})
    first = run("first", {"SF_ORDER_CACHE": str(cache)})
    files = list(cache.iterdir())
    assert len(files) == 1 and files[0].stat().st_size == 8 * n
    second = run("second", {"SF_ORDER_CACHE": str(cache)})
    assert plain == first == second
    # the cached ordering is what the engine's own analysis calls "Order"
    S = sf.analyze(sf.compress(n, r, c, x, sym), "lu")
    assert np.array_equal(np.fromfile(files[0], dtype=np.int64), S.array("Order"))
    files[0].write_bytes(b"\0" * (8 * n))          # not a permutation: must fall back to METIS and repair the file
    assert run("third", {"SF_ORDER_CACHE": str(cache)}) == plain
    assert np.array_equal(np.fromfile(files[0], dtype=np.int64), S.array("Order"))


def test_analyze_ordered_accepts_any_permutation():
    """A caller-supplied ordering (here: identity and a reversal) gives a valid structure whose CPU-interpreted
    factorization has a tiny backward error; a non-permutation is refused."""
    n, r, c, x, sym = sf.matrices.generate("lap2d", 12)
    A = sf.compress(n, r, c, x, sym)
    for order in (np.arange(n), np.arange(n)[::-1].copy()):
        S = sf.analyze(A, "cholesky", order=order)
        assert S.double("t_order") == 0.0
        perm = S.array("Perm")
        assert np.array_equal(np.sort(perm), np.arange(n))
        assert S.array("Super")[-1] == n and S.int("xsize") == S.array("Lsxp")[-1]
    with pytest.raises(RuntimeError, match="not a permutation"):
        sf.analyze(A, "cholesky", order=np.zeros(n, dtype=np.int64))


def test_decimal_scanner_is_bit_exact_on_many_formats(probe, tmp_path):
    """Values written in every format a MatrixMarket writer might use -- %.17g, %.3f, %e, %.10E, integers, leading '.',
    trailing '.', '+' signs, huge / tiny exponents, more than 19 significant digits -- must come out with exactly the
    bits Python's (correctly rounded) float() gives for the same text, whether the scanner's fast path or strtod took it."""
    rng = np.random.default_rng(123)
    n = 20000
    mags = 10.0 ** rng.uniform(-30, 30, n) * rng.choice([-1.0, 1.0], n)
    texts = []
    for k, v in enumerate(mags):
        f = k % 12
        if f == 0: t = "%.17g" % v
        elif f == 1: t = "%.3f" % (v % 1000.0)
        elif f == 2: t = "%e" % v
        elif f == 3: t = "%.10E" % v
        elif f == 4: t = str(int(v % 1e9) + 1)
        elif f == 5: t = ("%.6f" % abs(v % 1.0 + 0.001)).lstrip("0")              # ".123456"
        elif f == 6: t = "%d." % (int(abs(v)) % 100000 + 1)                        # "123."
        elif f == 7: t = "+%.5g" % abs(v)
        elif f == 8: t = "%.25g" % v                                               # > 19 significant digits
        elif f == 9: t = "%.15g" % v
        elif f == 10: t = "%.12ge%+d" % (rng.uniform(1, 10), int(rng.integers(-300, 300)))
        else: t = "%.16g" % v
        if float(t) == 0.0:
            t = "1"
        texts.append(t)
    texts[7] = "1e-320"          # subnormal
    texts[19] = "1.7976931348623157e308"
    texts[31] = "4.9406564584124654e-324"
    path = tmp_path / "fmt.mtx"
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d 1 %d\n" % (n, n))
        for k, t in enumerate(texts):
            f.write("%d 1 %s\n" % (k + 1, t))
    for threads in (1, 4):
        got, out = probe(path, threads)
        assert got is not None, out
        assert got["nz"] == n and np.array_equal(got["Ci"], np.arange(n))
        want = np.array([float(t) for t in texts])
        bad = np.flatnonzero(got["Cx"].view(np.int64) != want.view(np.int64))
        assert bad.size == 0, [(texts[i], got["Cx"][i], want[i]) for i in bad[:5]]


def greedy_static_pivot(n, Cp, Ci, Cx):
    """The reference's SparseFrame_pivot (LU.c:589-673) restated: column by column, the unused row of largest modulus."""
    piv = -np.ones(n, dtype=np.int64)
    for j in range(n):
        best, big = -1, 0.0
        for p in range(Cp[j], Cp[j + 1]):
            a = abs(Cx[p]) ** 2
            if piv[Ci[p]] < 0 and a > big:
                best, big = Ci[p], a
        if best >= 0:
            piv[best] = j
    free = iter(sorted(set(range(n)) - set(piv[piv >= 0].tolist())))
    for i in range(n):
        if piv[i] < 0:
            piv[i] = next(free)
    return piv


@pytest.mark.parametrize("cplx", [False, True])
def test_static_pivoting_prepass(probe, tmp_path, cplx):
    """SF_STATIC_PIVOT=1 runs the reference's (compiled-out, LU.c:784-787) SparseFrame_pivot after compress: rows are
    relabelled so that a large entry of every column sits on the diagonal.  A convection-diffusion matrix whose rows were
    rotated by one position (its diagonal is then a weak off-diagonal entry, or empty) comes back as the original, diagonally
    dominant matrix; what is then factorized equals the reference's factorization of the same relabelled triplets."""
    import sforacle
    import sfref
    if cplx:
        n, r, c, x, sym = sf.matrices.generate_complex("cd3d", 6, hermitian=False)
    else:
        n, r, c, x, sym = sf.matrices.generate("cd3d", 6)
    shifted = (r + 1) % n
    mtx = tmp_path / "shifted.mtx"
    sf.matrices.write_matrix_market(str(mtx), n, shifted, c, x, sym)
    plain, out = probe(mtx)
    assert plain is not None, out
    os.environ["SF_STATIC_PIVOT"] = "1"
    try:
        got, out = probe(mtx)
    finally:
        del os.environ["SF_STATIC_PIVOT"]
    assert got is not None, out
    cx = plain["Cx"].view(np.complex128) if cplx else plain["Cx"]
    piv = greedy_static_pivot(n, plain["Cp"], plain["Ci"], cx)
    assert np.array_equal(np.sort(piv), np.arange(n))
    assert np.array_equal(got["Ci"], piv[plain["Ci"]]) and np.array_equal(got["Cp"], plain["Cp"])
    A0 = sf.compress(n, r, c, x, sym)
    assert np.array_equal(got["Ci"], A0.Ci)              # the rotation is undone: the dominant diagonal is back
    check = sforacle.check_plan_complex if cplx else sforacle.plan_check
    res = check("lu", n, got["Cp"], got["Ci"], cx, False, sf.DEFAULT_DEVSLOT)
    assert res[1] == 0
    if sfref.available("lu"):
        cols = np.repeat(np.arange(n), np.diff(got["Cp"]))
        R = sfref.Reference(n, got["Ci"], cols, cx, False, "lu").analyze().factorize()
        ref = R.array("Lsx")
        assert np.abs(res[0] - ref).max() <= 1e-13 * np.abs(ref).max()


def test_analyze_refuses_a_rectangular_matrix(tmp_path):
    """The reader takes any shape (CH.c:449-521), the analysis needs a square matrix: the reference indexes Cp with row
    numbers regardless; the drop-in says so and returns."""
    src = tmp_path / "probe.c"
    src.write_text(ANALYZE_PROBE)
    exe = tmp_path / "probe"
    lib = os.path.join(PKG, "LU", "Lib")
    subprocess.check_call(["gcc", "-fopenmp", "-I" + os.path.join(PKG, "LU", "Include"), str(src), "-o", str(exe),
                           "-L" + lib, "-lSparseFrame", "-Wl,-rpath," + lib])
    mtx = tmp_path / "rect.mtx"
    mtx.write_text("%%MatrixMarket matrix coordinate real general\n3 2 2\n1 1 1.0\n3 2 2.0\n")
    p = subprocess.run([str(exe), str(mtx), str(tmp_path / "out")], capture_output=True, text=True)
    assert p.returncode == 4 and "not square" in p.stdout


def test_static_pivoting_is_the_identity_on_the_cfg4_generator(probe, tmp_path):
    """BASELINE config 4 is "LU ... (static pivoting)": on the convection-diffusion matrices the diagonal entry is the largest
    of its column, so the reference's pre-pass (SF_STATIC_PIVOT=1) relabels nothing and the factorization is the one measured."""
    n, r, c, x, sym = sf.matrices.generate("cd3d", 8)
    mtx = tmp_path / "cd.mtx"
    sf.matrices.write_matrix_market(str(mtx), n, r, c, x, sym)
    plain, out = probe(mtx)
    os.environ["SF_STATIC_PIVOT"] = "1"
    try:
        piv, out = probe(mtx)
    finally:
        del os.environ["SF_STATIC_PIVOT"]
    assert plain is not None and piv is not None
    assert np.array_equal(plain["Ci"], piv["Ci"]) and np.array_equal(plain["Cp"], piv["Cp"])


@pytest.mark.parametrize("seed", range(int(os.environ.get("SF_READER_FUZZ", "12"))))
def test_reader_against_the_references_own_reader(probe, tmp_path, seed):
    """Random well-formed MatrixMarket files -- number formats, signs, exponents, tabs and leading blanks, comment lines, a
    fourth column in real files, explicit zeros, every banner word -- read by the drop-in and by the REFERENCE'S OWN reader
    (SparseFrame_read_matrix_triplet + SparseFrame_compress, CH.c:400-587, through oracle/ref_driver.c): the same flags and
    the same Cp / Ci / Cx bit for bit."""
    import sfref
    if not sfref.available("lu"):
        pytest.skip("compiled reference not built")
    rng = np.random.default_rng(100 + seed)
    field = str(rng.choice(["real", "real", "complex", "integer"]))        # everything but "real" is complex to the reader
    symm = str(rng.choice(["general", "symmetric", "hermitian", "skew-symmetric"]))
    cplx = field != "real"
    n = int(rng.integers(3, 40))
    nz = int(rng.integers(1, 6 * n))
    ti, tj = rng.integers(1, n + 1, nz), rng.integers(1, n + 1, nz)

    def number(v):
        f = int(rng.integers(0, 9))
        if f == 0: return "%.17g" % v
        if f == 1: return "%.3f" % v
        if f == 2: return "%e" % v
        if f == 3: return "%.10E" % v
        if f == 4: return "%d" % int(v * 10)
        if f == 5: return ("+%.5g" % abs(v))
        if f == 6: return "%.25g" % v
        if f == 7: return "0" if rng.random() < 0.5 else "0.0e0"           # explicit zero
        return "%.12ge%+d" % (rng.uniform(1, 10), int(rng.integers(-30, 30)))
    lines = ["%%%%MatrixMarket matrix coordinate %s %s" % (field, symm)]
    for _ in range(int(rng.integers(0, 4))):
        lines.append("%% comment %d" % rng.integers(0, 100))
    lines.append("%d %d %d" % (n, n, nz))
    for k in range(nz):
        v, w = rng.standard_normal() * 10.0 ** rng.integers(-5, 6), rng.standard_normal()
        sep = str(rng.choice([" ", "  ", "\t", " \t "]))
        lead = str(rng.choice(["", " ", "\t"]))
        parts = [str(ti[k]), str(tj[k]), number(v)]
        if cplx or rng.random() < 0.2:
            parts.append(number(w))                                          # a 4th column in a real file is legal (CH.c:485-496)
        lines.append(lead + sep.join(parts))
    path = tmp_path / "r.mtx"
    path.write_text("\n".join(lines) + "\n")
    R = sfref.Reference.from_file(path, "lu")
    got, out = probe(path, int(rng.integers(1, 4)))
    assert got is not None, out
    assert (got["nrow"], got["ncol"], got["nz"], got["sym"], got["cplx"]) == \
        (R.n, R.n, R.long("nzmax"), R.long("isSymmetric"), R.long("isComplex"))
    assert np.array_equal(got["Cp"], R.array("Cp")) and np.array_equal(got["Ci"], R.array("Ci"))
    ref_cx = R.array("Cx")
    assert np.array_equal(got["Cx"].view(np.int64), ref_cx.view(np.float64).view(np.int64))
