import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparse-matrix-factorization-library_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load_golden(path):
    d = dict(np.load(path))
    gen, fk, symtxt = [str(v) for v in d["meta"]]
    n, ns, isize, xsize, csize, nstage, nsleaf, slot, g = [int(v) for v in d["scalars"]]
    d.update(gen=gen, kind=fk, symmetric=(symtxt == "symmetric"), n=n, nsuper=ns, isize=isize, xsize=xsize,
             csize=csize, nstage=nstage, nsleaf=nsleaf, devslot=slot, grid=g, name=os.path.basename(path)[:-4])
    return d


@pytest.fixture(params=GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def golden(request):
    return load_golden(request.param)


def build_matrix(gen, g):
    import sparseframe_b200 as sf
    n, r, c, x, sym = sf.matrices.generate(gen, g)
    return sf.compress(n, r, c, x, sym), (n, r, c, x, sym)
