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
                "SparseFrame_solve_supernodal", "SparseFrame_validate", "SparseFrame_cleanup_matrix"]

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
