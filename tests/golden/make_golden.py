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

INT_ARRAYS = ["Perm", "Parent", "Post", "ColCount", "Lp", "Li", "LTp", "LTi", "Super", "SuperMap", "Sparent", "Lsip",
              "Lsxp", "Lsi", "ST_Map", "ST_Pointer", "ST_Index", "LeafQueue", "Aoffset", "Moffset"]


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, gen, g, fk, slot in CASES:
        n, r, c, x, sym = matrices.generate(gen, g)
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
        d["residual"] = np.array([ref.validate()])
        d["scalars"] = np.array([n, ns, ref.long("isize"), ref.long("xsize"), ref.long("csize"), ref.long("nstage"),
                                 ref.long("nsleaf"), slot, g], dtype=np.int64)
        d["meta"] = np.array([gen, fk, "symmetric" if sym else "general"])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "n", n, "nsuper", ns, "xsize", ref.long("xsize"), "residual %.2e" % d["residual"][0])
        ref.close()


if __name__ == "__main__":
    main()
