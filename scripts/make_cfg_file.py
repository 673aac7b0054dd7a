"""Writes a BASELINE config's matrix as the drop-in's binary triplet container (reads in a fraction of a second):
    python scripts/make_cfg_file.py cfg2 /tmp/cfg2.sfb"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
from sparseframe_b200 import matrices  # noqa: E402

cfg, path = sys.argv[1], sys.argv[2]
gen, g, kind = matrices.CONFIGS[cfg] if cfg in matrices.CONFIGS else (cfg.split(":")[0], int(cfg.split(":")[1]), "cholesky")
matrices.write_binary(path, *matrices.generate(gen, g))
print(cfg, gen, g, kind, "->", path)
