"""sparseframe_b200 -- host-side mirror of the SparseFrame interface over the B200 numeric engine.

The product is the native library (``lib/libsfb200.so``, C ABI in ``include/sfb200.h``) and the drop-in
``Cholesky/`` / ``LU/`` C layout next to this package.  This Python layer exists for the tests and the benchmark:
it follows the reference's call sequence -- ``read`` (here: triplets -> ``compress``), ``analyze``, ``factorize``,
``solve_supernodal``, ``validate`` (``Cholesky/Source/SparseFrame.c:3390-3434``) -- with the same names and meaning.
"""
from . import matrices  # noqa: F401
from .api import (DEFAULT_DEVSLOT, Matrix, Symbolic, Plan, ZPlan, ZGroup, Group, wrap_structure, compress, analyze, factorize, solve_supernodal, validate,
                  lower_trapezoid_mask, entrywise_rel_diff, device_count, fp64_peak)  # noqa: F401
from ._capi import LIB_PATH, SYMBOLS, LibraryMissing  # noqa: F401
