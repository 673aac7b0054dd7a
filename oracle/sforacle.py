"""oracle/sforacle.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libsforacle.so (oracle/sf_oracle.c, the
plain-C restatement of the reference's numeric path) and of oracle/_ref/libsfcheck.so (oracle/plan_host_exec.cpp, the
CPU interpreter of the engine's schedule).  Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may
import this module, and only as the checker."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_chk = None


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(_HERE, "_ref", "libsforacle.so"))
        _lib.sforacle_factorize.restype = C.c_long
        _lib.sforacle_factorize.argtypes = [C.c_int, C.c_long, C.c_long] + [C.c_void_p] * 5 + [C.c_long, C.c_long] + \
            [C.c_void_p] * 8
        _lib.sforacle_solve.restype = None
        _lib.sforacle_solve.argtypes = [C.c_int, C.c_long, C.c_long] + [C.c_void_p] * 7
        _lib.sforacle_residual.restype = C.c_double
        _lib.sforacle_residual.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 9
    return _lib


def factorize(kind, st, Lx, Ux=None):
    """st: dict with Super, SuperMap, Lsip, Lsxp, Lsi, LeafQueue, Lp, Li (+Up, Ui), scalars csize, nsleaf."""
    lu = kind in (1, "lu")
    n = len(st["SuperMap"])
    ns = len(st["Super"]) - 1
    arr = {k: _i64(st[k]) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    Lx = np.ascontiguousarray(Lx, dtype=np.float64)
    up = ui = ux = None
    if lu and Ux is not None:
        up, ui, ux = _i64(st["Up"]), _i64(st["Ui"]), np.ascontiguousarray(Ux, dtype=np.float64)
    out = np.zeros(int(arr["Lsxp"][-1]), dtype=np.float64)
    bad = lib().sforacle_factorize(
        int(lu), n, ns, arr["Super"].ctypes.data, arr["SuperMap"].ctypes.data, arr["Lsip"].ctypes.data,
        arr["Lsxp"].ctypes.data, arr["Lsi"].ctypes.data, int(st["csize"]), int(st["nsleaf"]),
        arr["LeafQueue"].ctypes.data, arr["Lp"].ctypes.data, arr["Li"].ctypes.data, Lx.ctypes.data,
        up.ctypes.data if up is not None else None, ui.ctypes.data if ui is not None else None,
        ux.ctypes.data if ux is not None else None, out.ctypes.data)
    return out, int(bad)


def solve(kind, st, Lsx, b):
    lu = kind in (1, "lu")
    n = len(b)
    ns = len(st["Super"]) - 1
    arr = {k: _i64(st[k]) for k in ("Super", "Lsip", "Lsxp", "Lsi")}
    Lsx = np.ascontiguousarray(Lsx, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty(n, dtype=np.float64)
    lib().sforacle_solve(int(lu), n, ns, arr["Super"].ctypes.data, arr["Lsip"].ctypes.data, arr["Lsxp"].ctypes.data,
                         arr["Lsi"].ctypes.data, Lsx.ctypes.data, b.ctypes.data, x.ctypes.data)
    return x


def residual(kind, st, Lx, b, x, Ux=None):
    lu = kind in (1, "lu")
    n = len(b)
    lp, li = _i64(st["Lp"]), _i64(st["Li"])
    Lx = np.ascontiguousarray(Lx, dtype=np.float64)
    up = ui = ux = None
    if lu and Ux is not None:
        up, ui, ux = _i64(st["Up"]), _i64(st["Ui"]), np.ascontiguousarray(Ux, dtype=np.float64)
    out = np.zeros(4)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    r = lib().sforacle_residual(int(lu), n, lp.ctypes.data, li.ctypes.data, Lx.ctypes.data,
                                up.ctypes.data if up is not None else None, ui.ctypes.data if ui is not None else None,
                                ux.ctypes.data if ux is not None else None, b.ctypes.data, x.ctypes.data,
                                out.ctypes.data)
    return float(r), out


def validate(kind, st, Lx, Lsx, Ux=None):
    """reference self-check: b_i = 1 + i/n (CH.c:3186) -> (printed residual, backward error ||r|| / (||A|| ||x||))"""
    n = len(st["SuperMap"])
    b = 1.0 + np.arange(n, dtype=np.float64) / n
    x = solve(kind, st, Lsx, b)
    res, norms = residual(kind, st, Lx, b, x, Ux)
    return res, float(norms[3] / (norms[0] * norms[2]))


def plan_check(kind, A_n, Cp, Ci, Cx, symmetric, devslot, nb=0, outer=0, small_a=-1, small_b=-1, ld_align=0):
    """analyze + build the engine's schedule + interpret it on the CPU (oracle/plan_host_exec.cpp)."""
    global _chk
    if _chk is None:
        _chk = C.CDLL(os.path.join(_HERE, "_ref", "libsfcheck.so"))
        _chk.sfcheck_factorize.restype = C.c_long
        _chk.sfcheck_factorize.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_long]
    Cp, Ci = _i64(Cp), _i64(Ci)
    Cx = np.ascontiguousarray(Cx, dtype=np.float64)
    stats = np.zeros(8, dtype=np.int64)
    # first call sizes the output
    rc = _chk.sfcheck_factorize(int(kind in (1, "lu")), int(symmetric), int(A_n), Cp.ctypes.data, Ci.ctypes.data,
                                Cx.ctypes.data, int(devslot), nb, outer, small_a, small_b, ld_align, None,
                                stats.ctypes.data, 0)
    if rc < 0:
        raise RuntimeError("plan check failed: %d" % rc)
    out = np.zeros(int(stats[6]), dtype=np.float64)
    rc = _chk.sfcheck_factorize(int(kind in (1, "lu")), int(symmetric), int(A_n), Cp.ctypes.data, Ci.ctypes.data,
                                Cx.ctypes.data, int(devslot), nb, outer, small_a, small_b, ld_align, out.ctypes.data,
                                stats.ctypes.data, out.size)
    return out, int(rc), stats


class CheckBackend:
    """One rank of a `world`-way etree partition executed by the CPU plan interpreter (oracle/plan_host_exec.cpp).
    Same backend protocol as sparseframe_b200.parallel.DeviceBackend, so that the multi-rank orchestration can be
    tested with gloo on a box without GPUs."""

    def __init__(self, kind, A, devslot, rank, world, nb=0, outer=0, small_a=-1, small_b=-1):
        import torch
        global _chk
        if _chk is None:
            _chk = C.CDLL(os.path.join(_HERE, "_ref", "libsfcheck.so"))
        L = _chk
        L.sfcheck_open.restype = C.c_void_p
        L.sfcheck_open.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + [C.c_int] * 7
        L.sfcheck_int.restype = C.c_long
        L.sfcheck_int.argtypes = [C.c_void_p, C.c_char_p]
        L.sfcheck_double.restype = C.c_double
        L.sfcheck_double.argtypes = [C.c_void_p, C.c_char_p]
        for f in ("sfcheck_segment", "sfcheck_coll"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfcheck_owner.argtypes = [C.c_void_p, C.c_void_p]
        L.sfcheck_begin.argtypes = [C.c_void_p, C.c_void_p]
        L.sfcheck_run_segment.restype = C.c_long
        L.sfcheck_run_segment.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfcheck_pack_owned.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sfcheck_close.argtypes = [C.c_void_p]
        self.L = L
        self.A = A
        self.h = L.sfcheck_open(int(kind in (1, "lu")), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data,
                                A.Cx.ctypes.data, int(devslot), nb, outer, small_a, small_b, 0, rank, world)
        if not self.h:
            raise RuntimeError("sfcheck_open failed")
        self.rank, self.world = rank, world
        self.X = np.zeros(self.int("xsize_work"))      # factor image, then this rank's scratch area
        self.image = torch.from_numpy(self.X)[:self.int("xsize_pad")]          # shares memory with X
        self.bad = 0

        def info(fn, k, n):
            out = np.zeros(n, dtype=np.int64)
            fn(self.h, k, out.ctypes.data)
            return [int(v) for v in out]
        self.segments = [dict(zip(("coll_first", "coll_count", "launch_first", "launch_count", "level"),
                                  info(L.sfcheck_segment, k, 5))) for k in range(self.int("nsegments"))]
        self.colls = [dict(zip(("type", "offset", "count", "root", "lo", "hi"), info(L.sfcheck_coll, i, 6)))
                      for i in range(self.int("ncolls"))]

    def int(self, name):
        return int(self.L.sfcheck_int(self.h, name.encode()))

    def double(self, name):
        return float(self.L.sfcheck_double(self.h, name.encode()))

    def owner(self):
        out = np.zeros(self.int("nsuper"), dtype=np.int32)
        self.L.sfcheck_owner(self.h, out.ctypes.data)
        return out

    def begin(self):
        self.X[:] = 0.0
        self.bad = 0
        self.L.sfcheck_begin(self.h, self.X.ctypes.data)

    def run_segment(self, k):
        self.bad += self.L.sfcheck_run_segment(self.h, k, self.X.ctypes.data)

    def finish(self):
        pass

    def pack_owned(self):
        out = np.zeros(self.int("xsize"))
        self.L.sfcheck_pack_owned(self.h, self.X.ctypes.data, out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.L.sfcheck_close(self.h)
            self.h = None
