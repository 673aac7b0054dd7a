"""oracle/sfref.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libsfref_{chol,lu}.so.

Those libraries are the UNMODIFIED reference CPU path (compiled from /root/reference by
oracle/Makefile) behind oracle/ref_driver.c.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the checker / baseline.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

# reference's devSlotSize for one B200 (GPU_SPLIT_LIMIT=4 -> 4 virtual GPUs, 8 slots each; CH.c:81-87,199),
# the value SURVEY.md §8 measured the structures with.  Pinned so that symbolic structure is well defined.
DEFAULT_DEVSLOT = 5_380_000_000


def available(kind: str = "cholesky") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libsfref_%s.so" % ("lu" if kind == "lu" else "chol")))


def _lib(kind: str):
    key = "lu" if kind == "lu" else "chol"
    if key not in _LIBS:
        lib = C.CDLL(os.path.join(_HERE, "_ref", "libsfref_%s.so" % key))
        lib.sfref_create.restype = C.c_void_p
        lib.sfref_create.argtypes = [C.c_int, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_ulonglong, C.c_int]
        lib.sfref_create_from_file.restype = C.c_void_p
        lib.sfref_create_from_file.argtypes = [C.c_char_p, C.c_ulonglong, C.c_int]
        lib.sfref_create_complex.restype = C.c_void_p
        lib.sfref_create_complex.argtypes = lib.sfref_create.argtypes
        for f in ("sfref_analyze", "sfref_factorize", "sfref_validate"):
            getattr(lib, f).restype = C.c_int
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.sfref_long.restype = C.c_long
        lib.sfref_long.argtypes = [C.c_void_p, C.c_char_p]
        lib.sfref_double.restype = C.c_double
        lib.sfref_double.argtypes = [C.c_void_p, C.c_char_p]
        lib.sfref_array.restype = C.c_void_p
        lib.sfref_array.argtypes = [C.c_void_p, C.c_char_p]
        lib.sfref_destroy.restype = None
        lib.sfref_destroy.argtypes = [C.c_void_p]
        _LIBS[key] = lib
    return _LIBS[key]


class Reference:
    """One matrix going through the reference's analyze / factorize / validate.  Complex `vals` select the reference's
    isComplex path (values travel as interleaved (re, im) pairs, value arrays come back as complex128)."""

    def __init__(self, n, rows, cols, vals, symmetric, kind="cholesky", devslot=DEFAULT_DEVSLOT, blas_threads=0):
        self.kind = kind
        self.lib = _lib(kind)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        self.complex = np.iscomplexobj(vals)
        self.vdtype = np.complex128 if self.complex else np.float64
        vals = np.ascontiguousarray(vals, dtype=self.vdtype)
        create = self.lib.sfref_create_complex if self.complex else self.lib.sfref_create
        self.h = create(int(bool(symmetric)), int(n), int(rows.size), rows.ctypes.data,
                        cols.ctypes.data, vals.ctypes.data, int(devslot), int(blas_threads))
        self.n = int(n)

    @classmethod
    def from_file(cls, path, kind="cholesky", devslot=DEFAULT_DEVSLOT, blas_threads=0):
        """The matrix read by the reference's OWN MatrixMarket reader (SparseFrame_read_matrix, CH.c:652-691)."""
        self = cls.__new__(cls)
        self.kind = kind
        self.lib = _lib(kind)
        self.h = self.lib.sfref_create_from_file(str(path).encode(), int(devslot), int(blas_threads))
        if not self.h:
            raise RuntimeError("the reference could not read %s" % path)
        self.n = self.long("nrow")
        self.complex = self.long("isComplex") == 1
        self.vdtype = np.complex128 if self.complex else np.float64
        return self

    def analyze(self):
        self.lib.sfref_analyze(self.h)
        return self

    def factorize(self):
        self.lib.sfref_factorize(self.h)
        return self

    def validate(self):
        self.lib.sfref_validate(self.h)
        return self.double("residual")

    def long(self, name):
        return int(self.lib.sfref_long(self.h, name.encode()))

    def double(self, name):
        return float(self.lib.sfref_double(self.h, name.encode()))

    def _arr(self, name, count, dtype):
        p = self.lib.sfref_array(self.h, name.encode())
        if not p or count == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=count).copy()

    def array(self, name):
        n, ns = self.n, self.long("nsuper")
        nz = self.long("nzmax")
        sizes = {
            "Cp": n + 1, "Ci": nz, "Lp": n + 1, "LTp": n + 1, "Up": n + 1, "UTp": n + 1,
            "Perm": n, "Parent": n, "Post": n, "ColCount": n,
            "Super": ns + 1, "SuperMap": n, "Sparent": ns, "LeafQueue": ns,
            "Lsip": ns + 1, "Lsxp": ns + 1, "Lsi": self.long("isize"),
            "ST_Map": ns, "ST_Index": ns, "ST_Pointer": self.long("nstage") + 1,
            "Aoffset": ns, "Moffset": ns,
        }
        if name in ("Li", "LTi", "Ui", "UTi"):
            ptr = {"Li": "Lp", "LTi": "LTp", "Ui": "Up", "UTi": "UTp"}[name]
            return self._arr(name, int(self.array(ptr)[-1]), np.int64)
        if name in ("Lx", "LTx", "Ux", "UTx"):
            ptr = {"Lx": "Lp", "LTx": "LTp", "Ux": "Up", "UTx": "UTp"}[name]
            return self._arr(name, int(self.array(ptr)[-1]), self.vdtype)
        if name == "Cx":
            return self._arr(name, nz, self.vdtype)
        if name == "Lsx":
            return self._arr(name, self.long("xsize"), self.vdtype)
        if name in ("Xx", "Bx"):
            return self._arr(name, n, self.vdtype)
        return self._arr(name, sizes[name], np.int64)

    def structure(self):
        return {k: self.array(k) for k in ("Perm", "Super", "SuperMap", "Sparent", "Lsip", "Lsxp", "Lsi")}

    def close(self):
        if self.h:
            self.lib.sfref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
