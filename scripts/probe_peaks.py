import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
import sparseframe_b200 as sf
for which in ("dfma", "dmma", "dmma_reg_tile", "dmma_smem_tile"):
    for it in (2048, 8192):
        print("%-16s iters %5d: %.2f TFLOP/s" % (which, it, sf.fp64_peak(0, which, it)), flush=True)
