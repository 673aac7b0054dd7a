"""ctypes mirror of the drop-in C library (Cholesky/Lib/libSparseFrame.so, LU/Lib/libSparseFrame.so): the three structs of
Include/info.h and the reference's entry points, so that Python callers -- bench.py's end-to-end leg, the tests -- go
through the very symbols a C program linked against the reference would call (SparseFrame_factorize, CH.c:3019 /
LU.c:3575).  Member order and types are the ABI (reference Include/info.h:12-150); tests/test_dropin.py checks sizeof /
offsetof of these classes against the numbers measured on the reference's headers."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

Long = C.c_long
Float = C.c_double
PLong = C.POINTER(C.c_long)
PFloat = C.POINTER(C.c_double)
PSize = C.POINTER(C.c_size_t)

TYPE_CHOLESKY, TYPE_QR, TYPE_LU = 0, 1, 2
PERM_IDENTITY, PERM_AMD, PERM_METIS = 0, 1, 2


class common_info_struct(C.Structure):
    _fields_ = [("numCPU", C.c_int), ("numGPU", C.c_int), ("numGPU_physical", C.c_int), ("minDevMemSize", C.c_size_t),
                ("minHostMemSize", C.c_size_t), ("matrixThreadNum", C.c_int), ("numSparseMatrix", C.c_int),
                ("devSlotSize", C.c_size_t), ("allocateTime", C.c_double), ("computeTime", C.c_double),
                ("freeTime", C.c_double)]


class gpu_info_struct(C.Structure):
    _fields_ = [("gpuIndex_physical", C.c_int), ("gpuLock", C.c_ubyte * 4), ("devMem", C.c_void_p), ("devMemSize", C.c_size_t),
                ("hostMem", C.c_void_p), ("hostMemSize", C.c_size_t), ("sharedMemSize", C.c_size_t),
                ("h_A", C.c_void_p * 2), ("h_B", C.c_void_p * 2), ("h_Lsi", C.c_void_p), ("d_A", C.c_void_p * 3),
                ("d_B", C.c_void_p * 2), ("d_C", C.c_void_p * 2), ("d_Lsi", C.c_void_p),
                ("s_cudaEvent_factorized", C.c_void_p), ("d_cudaEvent_onDevice", C.c_void_p * 2),
                ("d_cudaEvent_applied", C.c_void_p * 2), ("s_cudaStream", C.c_void_p), ("s_cudaStream_copyback", C.c_void_p),
                ("d_cudaStream", C.c_void_p * 2), ("d_cudaStream_copy", C.c_void_p * 2), ("s_cublasHandle", C.c_void_p),
                ("d_cublasHandle", C.c_void_p * 2), ("s_cusolverDnHandle", C.c_void_p), ("lastMatrix", C.c_int)]


def _matrix_fields(lu: bool):
    f = [("serial", C.c_int), ("path", C.c_char_p), ("file", C.c_void_p), ("factorizeType", C.c_int),
         ("isSymmetric", C.c_int), ("isComplex", C.c_int), ("ncol", Long), ("nrow", Long), ("nzmax", Long),
         ("Tj", PLong), ("Ti", PLong), ("Tx", PFloat), ("Cp", PLong), ("Ci", PLong), ("Cx", PFloat)]
    if lu:
        f += [("nzCPCT", Long), ("CPCTp", PLong), ("CPCTi", PLong)]
    f += [("Lp", PLong), ("Li", PLong), ("Lx", PFloat), ("LTp", PLong), ("LTi", PLong), ("LTx", PFloat)]
    if lu:
        f += [("Up", PLong), ("Ui", PLong), ("Ux", PFloat), ("UTp", PLong), ("UTi", PLong), ("UTx", PFloat)]
    f += [("permMethod", C.c_int)]
    if lu:
        f += [("PivInv", PLong)]
    f += [("Perm", PLong), ("Parent", PLong), ("Post", PLong), ("ColCount", PLong), ("nsuper", Long), ("Super", PLong),
          ("SuperMap", PLong), ("Sparent", PLong), ("nsleaf", Long), ("LeafQueue", PLong), ("isize", Long), ("xsize", Long),
          ("Lsip", PLong), ("Lsxp", PLong), ("Lsi", PLong), ("Lsx", PFloat), ("csize", Long), ("nstage", Long),
          ("ST_Map", PLong), ("ST_Pointer", PLong), ("ST_Index", PLong), ("ST_Parent", PLong), ("Aoffset", PSize),
          ("Moffset", PSize), ("workspace", C.c_void_p), ("workSize", C.c_size_t), ("Bx", PFloat), ("Xx", PFloat),
          ("Rx", PFloat), ("residual", Float), ("readTime", C.c_double), ("analyzeTime", C.c_double),
          ("factorizeTime", C.c_double), ("solveTime", C.c_double)]
    return f


class matrix_info_struct_cholesky(C.Structure):
    _fields_ = _matrix_fields(False)


class matrix_info_struct_lu(C.Structure):
    _fields_ = _matrix_fields(True)


class Library:
    """One of the two drop-in libraries with typed entry points."""

    def __init__(self, kind="cholesky"):
        self.lu = kind in ("lu", "LU", 1)
        tree = "LU" if self.lu else "Cholesky"
        self.path = os.path.join(_PKG, tree, "Lib", "libSparseFrame.so")
        if not os.path.exists(self.path):
            raise RuntimeError("%s is missing: run `make -C sparse-matrix-factorization-library_b200`" % self.path)
        self.lib = C.CDLL(self.path)
        self.matrix_info_struct = matrix_info_struct_lu if self.lu else matrix_info_struct_cholesky
        pc, pm = C.POINTER(common_info_struct), C.POINTER(self.matrix_info_struct)
        self.lib.SparseFrame_factorize.argtypes = [pc, C.POINTER(gpu_info_struct), pm]
        self.lib.SparseFrame_factorize.restype = C.c_int
        self.lib.SparseFrame_analyze.argtypes = [pc, pm]
        for name in ("SparseFrame_initialize_matrix", "SparseFrame_solve_supernodal", "SparseFrame_validate", "SparseFrame_cleanup_matrix"):
            getattr(self.lib, name).argtypes = [pm]
        self.lib.SparseFrame_time.restype = C.c_double

    def matrix_from_structure(self, S, lx, ux, lsx_ptr: int, keep: list):
        """matrix_info holding what SparseFrame_factorize reads (SURVEY.md 8b "Inputs"): the structure arrays of an analysis
        (copied into numpy arrays that `keep` holds alive), the matrix values and the output buffer.  lx / ux: numpy arrays or
        raw addresses (e.g. of page-locked buffers)."""
        import numpy as np
        m = self.matrix_info_struct()
        m.factorizeType = TYPE_LU if self.lu else TYPE_CHOLESKY
        m.permMethod = PERM_METIS
        m.isSymmetric = int(S.int("symmetric"))
        m.isComplex = int(S.int("complex") == 1)          # value arrays then hold Complex (interleaved pairs)
        m.nrow = m.ncol = S.int("n")
        m.nsuper, m.isize, m.xsize, m.csize = S.int("nsuper"), S.int("isize"), S.int("xsize"), S.int("csize")
        m.nsleaf, m.nstage = S.int("nsleaf"), S.int("nstage")

        def long_array(name):
            a = np.ascontiguousarray(S.array(name), dtype=np.int64)
            keep.append(a)
            return C.cast(a.ctypes.data, PLong)

        def addr(v):
            if isinstance(v, int):
                return C.cast(v, PFloat)
            keep.append(v)
            return C.cast(v.ctypes.data, PFloat)
        for name in ("Lp", "Li", "Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue"):
            setattr(m, name, long_array(name))
        m.nzmax = int(S.array("Lp")[-1])
        m.Lx = addr(lx)
        if self.lu and not m.isSymmetric:
            m.Up, m.Ui = long_array("Up"), long_array("Ui")
            m.Ux = addr(ux)
        m.Lsx = C.cast(lsx_ptr, PFloat)
        return m
