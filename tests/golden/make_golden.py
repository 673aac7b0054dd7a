"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libsfref_*.so, built from /root/reference by
oracle/Makefile).  Run in the build container (the GPU box has no /root/reference, but the prebuilt .so travels).

Each fixture holds, for one small deterministic matrix: the triplets are NOT stored (the generator is deterministic:
kind + grid), only the reference's outputs -- the full symbolic structure (bit-exact contract) and the factor Lsx
(1e-10 entrywise contract), plus the reference's printed residual.  devSlotSize is stored because the structure
depends on it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from sparseframe_b200 import matrices  # noqa: E402
import sfref  # noqa: E402

CASES = [
    # name, generator, grid, factorization, devSlotSize
    ("lap2d_20_chol", "lap2d", 20, "cholesky", sfref.DEFAULT_DEVSLOT),
    ("lap3d_8_chol", "lap3d", 8, "cholesky", sfref.DEFAULT_DEVSLOT),
    ("lap3d_10_chol_slot", "lap3d", 10, "cholesky", 40_000),          # tiny slot: capped supernodes, many stages
    ("el3d_4_chol", "el3d", 4, "cholesky", sfref.DEFAULT_DEVSLOT),
    ("cd3d_8_lu", "cd3d", 8, "lu", sfref.DEFAULT_DEVSLOT),
    ("lap2d_16_lu", "lap2d", 16, "lu", sfref.DEFAULT_DEVSLOT),        # LU on a symmetric input (U aliases L)
    ("cd3d_9_lu_slot", "cd3d", 9, "lu", 60_000),
]

# Complex matrices (matrix_info->isComplex): fixtures under tests/golden/complex/ (matrices.generate_complex; Hermitian for the
# zpotrf path, general phases for the LU path).  name, generator, grid, factorization, devSlotSize, diag_phase (-1: the
# generator's own phases on the diagonal too)
COMPLEX_CASES = [
    ("z_lap3d_7_chol", "lap3d", 7, "cholesky", sfref.DEFAULT_DEVSLOT, -1.0),
    ("z_lap2d_14_chol_slot", "lap2d", 14, "cholesky", 12_000, -1.0),  # slot caps counted in sizeof(Complex) (CH.c:1489-1496)
    ("z_el3d_3_chol", "el3d", 3, "cholesky", sfref.DEFAULT_DEVSLOT, -1.0),
    ("z_cd3d_7_lu", "cd3d", 7, "lu", sfref.DEFAULT_DEVSLOT, -1.0),
    ("z_lap2d_12_lu", "lap2d", 12, "lu", sfref.DEFAULT_DEVSLOT, -1.0),   # complex symmetric (A = A^T): U aliases L
    ("z_cd3d_6_lu_imagdiag", "cd3d", 6, "lu", 30_000, 1.5707963267948966),   # purely imaginary dominant diagonal
]

INT_ARRAYS = ["Perm", "Parent", "Post", "ColCount", "Lp", "Li", "LTp", "LTi", "Super", "SuperMap", "Sparent", "Lsip",
              "Lsxp", "Lsi", "ST_Map", "ST_Pointer", "ST_Index", "LeafQueue", "Aoffset", "Moffset"]


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    os.makedirs(os.path.join(out_dir, "complex"), exist_ok=True)
    cases = [c + (None,) for c in CASES] + [(os.path.join("complex", c[0]),) + c[1:] for c in COMPLEX_CASES]
    for name, gen, g, fk, slot, diag_phase in cases:
        if diag_phase is None:
            n, r, c, x, sym = matrices.generate(gen, g)
        else:
            n, r, c, x, sym = matrices.generate_complex(gen, g, hermitian=(fk == "cholesky"),
                                                        diag_phase=None if diag_phase < 0 else diag_phase)
        ref = sfref.Reference(n, r, c, x, sym, kind=fk, devslot=slot, blas_threads=1).analyze()
        leaf_queue = ref.array("LeafQueue")        # factorize appends to it (CH.c:2976): keep the analysis-time value
        ref.factorize()
        d = {k: ref.array(k) for k in INT_ARRAYS}
        d["LeafQueue_after"] = d["LeafQueue"]      # = the order in which the reference factorized the supernodes
        d["LeafQueue"] = leaf_queue
        ns = ref.long("nsuper")
        d["Super"] = d["Super"][:ns + 1]
        d["Sparent"] = d["Sparent"][:ns]
        d["Lx"] = ref.array("Lx")
        if fk == "lu" and not sym:
            for k in ("Up", "Ui", "UTp", "UTi"):
                d[k] = ref.array(k)
            d["Ux"] = ref.array("Ux")
        d["Lsx"] = ref.array("Lsx")
        # the reference's complex solve / validate are `// TODO` (CH.c:3088-3131): no residual to record
        d["residual"] = np.array([ref.validate() if diag_phase is None else np.nan])
        d["scalars"] = np.array([n, ns, ref.long("isize"), ref.long("xsize"), ref.long("csize"), ref.long("nstage"),
                                 ref.long("nsleaf"), slot, g], dtype=np.int64)
        d["meta"] = np.array([gen, fk, "symmetric" if sym else "general"])
        if diag_phase is not None:
            d["diag_phase"] = np.array([diag_phase])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "n", n, "nsuper", ns, "xsize", ref.long("xsize"), "residual %.2e" % d["residual"][0])
        ref.close()


if __name__ == "__main__":
    main()
