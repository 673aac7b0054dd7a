"""The drop-in libSparseFrame.so (Cholesky/ and LU/ trees): exported entry points and struct ABI."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-matrix-factorization-library_b200")

# exported entry points of the reference library: CH.c:3, 16, 287, 368, 385, 526, 589, 652, 1916, 3019,
# 3036, 3141, 3268, 3323
ENTRY_POINTS = ["SparseFrame", "SparseFrame_time", "SparseFrame_allocate_gpu", "SparseFrame_free_gpu",
                "SparseFrame_allocate_matrix", "SparseFrame_free_matrix", "SparseFrame_initialize_matrix",
                "SparseFrame_read_matrix", "SparseFrame_compress", "SparseFrame_analyze", "SparseFrame_factorize",
                "SparseFrame_solve_supernodal", "SparseFrame_validate", "SparseFrame_cleanup_matrix",
                # helper-kernel launchers of Include/cuda_kernel.cuh:9-23, kept as symbols (they launch nothing)
                "createMap", "createRelativeMap", "mappedSubtract", "deviceSum"]

# sizeof / offsetof measured on the REFERENCE headers (Include/info.h of each tree) with gcc x86-64:
# sizeof(common_info), sizeof(gpu_info), sizeof(matrix_info), then offsets of common.devSlotSize, gpu.lastMatrix,
# gpu.s_cublasHandle, matrix.Lp, matrix.Lsx, matrix.Perm, matrix.workspace, matrix.factorizeTime
REFERENCE_ABI = {
    "Cholesky": [72, 280, 440, 40, 272, 240, 112, 288, 168, 360, 424],
    "LU": [72, 280, 520, 40, 272, 240, 136, 368, 248, 440, 504],
}

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "SparseFrame.h"
int main(void) {
  printf("%zu %zu %zu ", sizeof(struct common_info_struct), sizeof(struct gpu_info_struct), sizeof(struct matrix_info_struct));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(struct common_info_struct, devSlotSize),
    offsetof(struct gpu_info_struct, lastMatrix), offsetof(struct gpu_info_struct, s_cublasHandle),
    offsetof(struct matrix_info_struct, Lp), offsetof(struct matrix_info_struct, Lsx), offsetof(struct matrix_info_struct, Perm),
    offsetof(struct matrix_info_struct, workspace), offsetof(struct matrix_info_struct, factorizeTime));
  return 0; }
"""


@pytest.mark.parametrize("tree", ["Cholesky", "LU"])
def test_entry_points_exported(tree):
    lib = C.CDLL(os.path.join(PKG, tree, "Lib", "libSparseFrame.so"))
    for name in ENTRY_POINTS:
        assert hasattr(lib, name), name


@pytest.mark.parametrize("tree", ["Cholesky", "LU"])
def test_struct_layout_matches_reference(tree, tmp_path):
    src = tmp_path / "probe.c"
    src.write_text(PROBE)
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-fopenmp", "-I" + os.path.join(PKG, tree, "Include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == REFERENCE_ABI[tree]


HEADERS = ["SparseFrame.h", "arch.h", "constant.h", "cuda_kernel.cuh", "extern.h", "info.h", "macro.h", "parameter.h", "type.h"]


@pytest.mark.parametrize("tree", ["Cholesky", "LU"])
def test_include_tree_has_the_reference_headers(tree, tmp_path):
    """north_star keeps the Include headers: every header of the reference's Include/ exists, compiles on its own as C,
    and parameter.h carries the policy functions with the reference's decisions (parameter.h:31-103)."""
    inc = os.path.join(PKG, tree, "Include")
    assert sorted(os.listdir(inc)) == sorted(HEADERS)
    for h in HEADERS:
        src = tmp_path / ("one_%s.c" % h.replace(".", "_"))
        src.write_text('#include "%s"\nint main(void) { return 0; }\n' % h)
        subprocess.check_call(["gcc", "-fopenmp", "-Wall", "-Wno-unused", "-I" + inc, str(src), "-o", str(tmp_path / "a.out")])
    src = tmp_path / "policy.c"
    src.write_text(r"""
#include <stdio.h>
#include "SparseFrame.h"
int main(void) {
  printf("%d %d %d %d %d ", should_relax(16, 0.9), should_relax(17, 0.9), should_relax(65, 0.11), should_relax(300, 0.04), should_relax(300, 0.06));
  printf("%ld %ld %ld ", get_node_score_from_dimension(16, 0, 16), get_node_score_from_dimension(32, 32, 32), get_node_score_from_dimension(20, 0, 20));
  printf("%d %d %d %ld %ld %d\n", set_factorize_location(255, 510), set_factorize_location(256, 256), MATRIX_THREAD_NUM, potrf_split_threshold, potrf_split_block, (int)minus_one[0]);
  return 0; }
""")
    exe = tmp_path / "policy"
    subprocess.check_call(["gcc", "-fopenmp", "-I" + inc, str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)], text=True).split() == "1 0 0 1 0 0 2048 -400 1 0 2 512 256 -1".split()


@pytest.mark.parametrize("tree", ["Cholesky", "LU"])
def test_python_mirror_of_the_structs_has_the_reference_layout(tree):
    """sparseframe_b200/dropin.py (what bench.py's end-to-end leg calls SparseFrame_factorize through) against the same
    numbers."""
    from sparseframe_b200 import dropin as D
    M = D.matrix_info_struct_lu if tree == "LU" else D.matrix_info_struct_cholesky
    got = [C.sizeof(D.common_info_struct), C.sizeof(D.gpu_info_struct), C.sizeof(M), D.common_info_struct.devSlotSize.offset,
           D.gpu_info_struct.lastMatrix.offset, D.gpu_info_struct.s_cublasHandle.offset, M.Lp.offset, M.Lsx.offset, M.Perm.offset,
           M.workspace.offset, M.factorizeTime.offset]
    assert got == REFERENCE_ABI[tree]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,gen,g", [("cholesky", "lap3d", 16), ("lu", "cd3d", 12)])
def test_factorize_symbol_called_like_a_c_program_would(kind, gen, g):
    """SparseFrame_factorize of the drop-in library through ctypes with a hand-filled matrix_info (structure from an analysis,
    page-locked Lx / Lsx): the path bench.py's e2e number takes."""
    import numpy as np
    import torch
    import sforacle
    import sparseframe_b200 as sf
    from sparseframe_b200 import dropin as D
    S = sf.analyze(sf.compress(*sf.matrices.generate(gen, g)), kind)
    lib = D.Library(kind)
    split = kind == "lu" and S.int("symmetric") == 0
    h_lx = torch.from_numpy(S.array("Lx")).pin_memory()
    h_ux = torch.from_numpy(S.array("Ux")).pin_memory() if split else None
    h_out = torch.zeros(S.int("xsize"), dtype=torch.float64).pin_memory()
    keep = []
    m = lib.matrix_from_structure(S, h_lx.data_ptr(), h_ux.data_ptr() if split else 0, h_out.data_ptr(), keep)
    common = D.common_info_struct(numGPU=1, numGPU_physical=1, devSlotSize=sf.DEFAULT_DEVSLOT)
    for _ in range(3):
        h_out.zero_()
        assert lib.lib.SparseFrame_factorize(C.byref(common), None, C.byref(m)) == 0
        assert m.factorizeTime > 0
    st = {k: S.array(k) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    st["csize"], st["nsleaf"] = S.int("csize"), S.int("nsleaf")
    ux = None
    if split:
        st["Up"], st["Ui"] = S.array("Up"), S.array("Ui")
        ux = S.array("Ux")
    ref, bad = sforacle.factorize(kind, st, S.array("Lx"), ux)
    mask = sf.lower_trapezoid_mask(kind, st["Super"], st["Lsip"], st["Lsxp"], len(ref))
    assert bad == 0 and sf.entrywise_rel_diff(ref[mask], h_out.numpy()[mask]) <= 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("tree,gen,g", [("Cholesky", "lap3d", 14), ("LU", "cd3d", 12)])
def test_demo_runs_and_reports_small_residual(tree, gen, g, tmp_path):
    import sparseframe_b200 as sf
    path = tmp_path / ("%s_%d.mtx" % (gen, g))
    sf.matrices.write_matrix_market(str(path), *sf.matrices.generate(gen, g))
    out = subprocess.check_output([os.path.join(PKG, tree, "Demo", "demo"), str(path)], text=True)
    line = [ln for ln in out.splitlines() if ln.startswith("residual")][0]
    assert float(line.split(":")[1]) < 1e-13, out
    assert "Factorize time:" in out


@pytest.mark.gpu
def test_driver_keeps_two_matrices_in_flight(tmp_path):
    """SparseFrame() processes its matrices MATRIX_THREAD_NUM = 2 at a time (CH.c:3371-3444): concurrent callers of
    SparseFrame_analyze / SparseFrame_factorize / SparseFrame_validate on different matrix_info and the same gpu_info_list."""
    import sparseframe_b200 as sf
    paths = []
    for gen, g in (("lap3d", 20), ("lap2d", 200), ("lap3d", 16), ("lap2d", 90), ("lap3d", 24)):
        p = tmp_path / ("%s_%d.mtx" % (gen, g))
        sf.matrices.write_matrix_market(str(p), *sf.matrices.generate(gen, g))
        paths.append(str(p))
    out = subprocess.check_output([os.path.join(PKG, "Cholesky", "Demo", "demo")] + paths, text=True)
    res = [float(ln.split(":")[1]) for ln in out.splitlines() if ln.startswith("residual")]
    names = [ln.split(":")[1].strip() for ln in out.splitlines() if ln.startswith("Matrix name")]
    assert len(res) == 5 and max(res) < 1e-13, out
    assert sorted(names) == sorted(os.path.basename(p) for p in paths)


# The reference calls its entry points from an OpenMP loop over matrices (MATRIX_THREAD_NUM concurrent callers, different
# matrix_info, CH.c:3375-3440): read + analyze must be re-entrant.  Two matrices analysed concurrently must come out
# exactly as when analysed one after the other (no device needed for these two steps).
CONCURRENT_PROBE = r"""
#include <stdio.h>
#include <string.h>
#include <omp.h>
#include "SparseFrame.h"
int SparseFrame_initialize_matrix(struct matrix_info_struct *);
int SparseFrame_read_matrix(struct matrix_info_struct *);
int SparseFrame_analyze(struct common_info_struct *, struct matrix_info_struct *);
int SparseFrame_cleanup_matrix(struct matrix_info_struct *);
static unsigned long long fnv(const void *p, size_t bytes, unsigned long long h) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h; }
static unsigned long long digest(struct matrix_info_struct *m) {
  unsigned long long h = 1469598103934665603ull;
  h = fnv(m->Perm, m->nrow * sizeof(long), h);     h = fnv(m->Super, (m->nsuper + 1) * sizeof(long), h);
  h = fnv(m->Lsip, (m->nsuper + 1) * sizeof(long), h); h = fnv(m->Lsxp, (m->nsuper + 1) * sizeof(long), h);
  h = fnv(m->Lsi, m->isize * sizeof(long), h);     h = fnv(m->Li, m->nzmax * sizeof(long), h);
  h = fnv(m->Lx, m->nzmax * sizeof(double), h);    h = fnv(m->ST_Map, m->nsuper * sizeof(long), h);
  return h; }
int main(int argc, char **argv) {
  static struct matrix_info_struct m[2];
  static struct common_info_struct c;
  unsigned long long seq[2], par[2];
  memset(&c, 0, sizeof(c)); c.devSlotSize = 5380000000ul;
  for (int pass = 0; pass < 2; ++pass) {
    #pragma omp parallel for num_threads(2) if(pass == 1)
    for (int k = 0; k < 2; ++k) {
      memset(&m[k], 0, sizeof(m[k])); m[k].path = argv[1 + k]; m[k].serial = k;
      SparseFrame_initialize_matrix(&m[k]);
      if (SparseFrame_read_matrix(&m[k]) || SparseFrame_analyze(&c, &m[k])) { (pass ? par : seq)[k] = 0; continue; }
      (pass ? par : seq)[k] = digest(&m[k]);
      SparseFrame_cleanup_matrix(&m[k]);
    }
  }
  printf("%llu %llu %llu %llu\n", seq[0], seq[1], par[0], par[1]);
  return 0; }
"""


@pytest.mark.parametrize("tree,gens", [("Cholesky", (("lap3d", 16), ("lap2d", 70))), ("LU", (("cd3d", 14), ("lap3d", 12)))])
def test_read_and_analyze_are_reentrant(tree, gens, tmp_path):
    import sparseframe_b200 as sf
    src = tmp_path / "probe.c"
    src.write_text(CONCURRENT_PROBE)
    exe = tmp_path / "probe"
    lib = os.path.join(PKG, tree, "Lib")
    subprocess.check_call(["gcc", "-fopenmp", "-I" + os.path.join(PKG, tree, "Include"), str(src), "-o", str(exe),
                           "-L" + lib, "-lSparseFrame", "-Wl,-rpath," + lib])
    paths = []
    for gen, g in gens:
        p = tmp_path / ("%s_%d.mtx" % (gen, g))
        sf.matrices.write_matrix_market(str(p), *sf.matrices.generate(gen, g))
        paths.append(str(p))
    for _ in range(3):
        out = subprocess.check_output([str(exe)] + paths, text=True).strip().splitlines()[-1].split()
        s0, s1, p0, p1 = (int(v) for v in out)
        assert s0 != 0 and s1 != 0 and s0 != s1
        assert (p0, p1) == (s0, s1)
