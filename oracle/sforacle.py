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
        # complex twins (oracle/sf_oracle_z.c)
        _lib.sforacle_zfactorize.restype = C.c_long
        _lib.sforacle_zfactorize.argtypes = _lib.sforacle_factorize.argtypes
        _lib.sforacle_zsolve.restype = None
        _lib.sforacle_zsolve.argtypes = _lib.sforacle_solve.argtypes
        _lib.sforacle_zresidual.restype = C.c_double
        _lib.sforacle_zresidual.argtypes = _lib.sforacle_residual.argtypes
    return _lib


def zfactorize(kind, st, Lz, Uz=None):
    """Complex twin of factorize (oracle/sf_oracle_z.c: the reference's isComplex branches); values are complex128."""
    lu = kind in (1, "lu")
    n = len(st["SuperMap"])
    ns = len(st["Super"]) - 1
    arr = {k: _i64(st[k]) for k in ("Super", "SuperMap", "Lsip", "Lsxp", "Lsi", "LeafQueue", "Lp", "Li")}
    Lz = np.ascontiguousarray(Lz, dtype=np.complex128)
    up = ui = uz = None
    if lu and Uz is not None:
        up, ui, uz = _i64(st["Up"]), _i64(st["Ui"]), np.ascontiguousarray(Uz, dtype=np.complex128)
    out = np.zeros(int(arr["Lsxp"][-1]), dtype=np.complex128)
    bad = lib().sforacle_zfactorize(
        int(lu), n, ns, arr["Super"].ctypes.data, arr["SuperMap"].ctypes.data, arr["Lsip"].ctypes.data,
        arr["Lsxp"].ctypes.data, arr["Lsi"].ctypes.data, int(st["csize"]), int(st["nsleaf"]),
        arr["LeafQueue"].ctypes.data, arr["Lp"].ctypes.data, arr["Li"].ctypes.data, Lz.ctypes.data,
        up.ctypes.data if up is not None else None, ui.ctypes.data if ui is not None else None,
        uz.ctypes.data if uz is not None else None, out.ctypes.data)
    return out, int(bad)


def zsolve(kind, st, Lsz, b):
    lu = kind in (1, "lu")
    n = len(b)
    ns = len(st["Super"]) - 1
    arr = {k: _i64(st[k]) for k in ("Super", "Lsip", "Lsxp", "Lsi")}
    Lsz = np.ascontiguousarray(Lsz, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.zeros(n, dtype=np.complex128)
    lib().sforacle_zsolve(int(lu), n, ns, arr["Super"].ctypes.data, arr["Lsip"].ctypes.data, arr["Lsxp"].ctypes.data,
                          arr["Lsi"].ctypes.data, Lsz.ctypes.data, b.ctypes.data, x.ctypes.data)
    return x


def zvalidate(kind, st, Lz, Lsz, Uz=None):
    """b as the reference sets it for a complex matrix (CH.c:3188-3192), solve, residual of the permuted system:
    (residual, norms[4], x)."""
    lu = kind in (1, "lu")
    n = len(st["SuperMap"])
    b = (1.0 + np.arange(n) / n) + 1j * ((n / 2 - np.arange(n)) / (3 * n))
    x = zsolve(kind, st, Lsz, b)
    Lp, Li = _i64(st["Lp"]), _i64(st["Li"])
    Lz = np.ascontiguousarray(Lz, dtype=np.complex128)
    up = ui = uz = None
    if lu and Uz is not None:
        up, ui, uz = _i64(st["Up"]), _i64(st["Ui"]), np.ascontiguousarray(Uz, dtype=np.complex128)
    norms = np.zeros(4)
    res = lib().sforacle_zresidual(int(lu), n, Lp.ctypes.data, Li.ctypes.data, Lz.ctypes.data,
                                   up.ctypes.data if up is not None else None, ui.ctypes.data if ui is not None else None,
                                   uz.ctypes.data if uz is not None else None, b.ctypes.data, x.ctypes.data, norms.ctypes.data)
    return float(res), norms, x


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


def _load_chk():
    """libsfcheck.so with the prototypes of its plain entry points (whoever loads it first sets them all)."""
    global _chk
    if _chk is None:
        _chk = C.CDLL(os.path.join(_HERE, "_ref", "libsfcheck.so"))
        _chk.sfcheck_factorize.restype = C.c_long
        _chk.sfcheck_factorize.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_long]
        _chk.sfcheck_factorize_complex.restype = C.c_long
        _chk.sfcheck_factorize_complex.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_ulonglong] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p]
    return _chk


def plan_check(kind, A_n, Cp, Ci, Cx, symmetric, devslot, nb=0, outer=0, small_a=-1, small_b=-1, ld_align=0):
    """analyze + build the engine's schedule + interpret it on the CPU (oracle/plan_host_exec.cpp)."""
    _load_chk()
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


def check_plan_complex(kind, A_n, Cp, Ci, Cz, symmetric, devslot, nb=0, outer=0, small_a=-1, small_b=-1, row_phase=False,
                       rhs=None):
    """A complex matrix on the CPU interpreter: analysis with the complex element size, real embedding
    (csrc/complex_embed.cpp), the ordinary schedule on the embedded structure, complex panels read back.
    Returns (Lsx complex128, failed pivots, stats, x or None)."""
    f = _load_chk().sfcheck_factorize_complex
    Cp, Ci = _i64(Cp), _i64(Ci)
    Cz = np.ascontiguousarray(Cz, dtype=np.complex128)
    stats = np.zeros(8, dtype=np.int64)
    args = (int(kind in (1, "lu")), int(symmetric), int(A_n), Cp.ctypes.data, Ci.ctypes.data, Cz.ctypes.data, int(devslot),
            nb, outer, small_a, small_b, int(bool(row_phase)))
    rc = f(*args, None, stats.ctypes.data, None)
    if rc < 0:
        raise RuntimeError("complex plan check failed: %d" % rc)
    out = np.zeros(int(stats[1]), dtype=np.complex128)
    x = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.complex128).copy()
    rc = f(*args, out.ctypes.data, stats.ctypes.data, None if x is None else x.ctypes.data)
    return out, int(rc), stats, x


def check_plan_complex_partitioned(kind, A_n, Cp, Ci, Cz, symmetric, devslot, world, nb=0, outer=0, small_a=-1, small_b=-1,
                                   row_phase=False, xsize=None):
    """A complex matrix on `world` ranks the way sfb200_zgroup_factorize composes it: partition of the embedded structure
    simulated in lock step, complex panels read out of the packed embedded factor.  Returns (Lsx complex128, failed pivots)."""
    f = _load_chk().sfcheck_factorize_complex_partitioned
    f.restype = C.c_long
    f.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + [C.c_int] * 6 + [C.c_void_p]
    Cp, Ci = _i64(Cp), _i64(Ci)
    Cz = np.ascontiguousarray(Cz, dtype=np.complex128)
    if xsize is None:
        xsize = check_plan_complex(kind, A_n, Cp, Ci, Cz, symmetric, devslot)[2][1]
    out = np.zeros(int(xsize), dtype=np.complex128)
    rc = f(int(kind in (1, "lu")), int(symmetric), int(A_n), Cp.ctypes.data, Ci.ctypes.data, Cz.ctypes.data, int(devslot),
           nb, outer, small_a, small_b, int(world), int(bool(row_phase)), out.ctypes.data)
    if rc < 0:
        raise RuntimeError("partitioned complex plan check failed: %d" % rc)
    return out, int(rc)


class CheckBackend:
    """One rank of a `world`-way etree partition executed by the CPU plan interpreter (oracle/plan_host_exec.cpp), behind
    the backend protocol of sparseframe_b200.parallel (exchange_handle / connect / factorize / finish) so that the
    multi-rank host logic can be tested with gloo on a box without GPUs.

    Stand-ins for the device mechanisms: the rank's image lives in a POSIX shared-memory segment (CUDA IPC mapping: the
    128-byte exchange handle carries the segment's name), remote update entries are subtracted directly in the peer's
    segment (red.global.add over NVLink), a pushed block is copied into the peers' segments (peer copies), and every
    flag wait is a gloo barrier of the step's group -- stricter than the pairwise flags, never weaker."""

    def __init__(self, kind, A, devslot, rank, world, nb=0, outer=0, small_a=-1, small_b=-1, dist=None):
        from multiprocessing import shared_memory
        L = _load_chk()
        L.sfcheck_open.restype = C.c_void_p
        L.sfcheck_open.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong] + [C.c_int] * 7
        L.sfcheck_int.restype = C.c_long
        L.sfcheck_int.argtypes = [C.c_void_p, C.c_char_p]
        L.sfcheck_double.restype = C.c_double
        L.sfcheck_double.argtypes = [C.c_void_p, C.c_char_p]
        L.sfcheck_segment.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfcheck_coll.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.sfcheck_image_elems.argtypes = [C.c_void_p, C.c_void_p]
        L.sfcheck_owner.argtypes = [C.c_void_p, C.c_void_p]
        L.sfcheck_begin.argtypes = [C.c_void_p, C.c_void_p]
        L.sfcheck_run_segment.restype = C.c_long
        L.sfcheck_run_segment.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfcheck_push.restype = C.c_long
        L.sfcheck_push.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfcheck_pack_owned.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sfcheck_close.argtypes = [C.c_void_p]
        self.L = L
        self.A = A
        self.dist = dist
        self.h = L.sfcheck_open(int(kind in (1, "lu")), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data,
                                A.Cx.ctypes.data, int(devslot), nb, outer, small_a, small_b, 0, rank, world)
        if not self.h:
            raise RuntimeError("sfcheck_open failed")
        self.rank, self.world = rank, world
        elems = max(self.int("xsize_work"), 1)          # factor image, then this rank's scratch area
        self.shm = shared_memory.SharedMemory(create=True, size=8 * elems)
        self.X = np.ndarray((elems,), dtype=np.float64, buffer=self.shm.buf)
        self.X[:] = 0.0
        self.peers_shm = [None] * world
        self.img = (C.c_void_p * world)()
        self.img[rank] = self.X.ctypes.data
        self.connected = world == 1
        self.bad = 0
        self.pushed_bytes = 0
        self.groups = {}

        def info(fn, k, n, *extra):
            out = np.zeros(n, dtype=np.int64)
            fn(self.h, k, out.ctypes.data, *extra)
            return [int(v) for v in out]
        self.segments = [dict(zip(("coll_first", "coll_count", "launch_first", "launch_count", "level"),
                                  info(L.sfcheck_segment, k, 5))) for k in range(self.int("nsegments"))]
        self.colls = [dict(zip(("type", "root", "lo", "hi", "offset", "rows", "cols", "ld"), info(L.sfcheck_coll, i, 8, None)))
                      for i in range(self.int("ncolls"))]
        image_elems = np.zeros(world, dtype=np.int64)
        L.sfcheck_image_elems(self.h, image_elems.ctypes.data)
        self.image_elems = [int(v) for v in image_elems]

    def int(self, name):
        return int(self.L.sfcheck_int(self.h, name.encode()))

    def double(self, name):
        return float(self.L.sfcheck_double(self.h, name.encode()))

    def owner(self):
        out = np.zeros(self.int("nsuper"), dtype=np.int32)
        self.L.sfcheck_owner(self.h, out.ctypes.data)
        return out

    # ---- backend protocol of sparseframe_b200.parallel ------------------------------------------------------------
    def exchange_handle(self) -> bytes:
        return self.shm.name.encode().ljust(128, b"\0")

    def connect(self, handles: bytes):
        from multiprocessing import shared_memory
        for q in range(self.world):
            if q == self.rank:
                continue
            name = handles[128 * q:128 * (q + 1)].rstrip(b"\0").decode()
            self.peers_shm[q] = shared_memory.SharedMemory(name=name)
            view = np.ndarray((self.peers_shm[q].size // 8,), dtype=np.float64, buffer=self.peers_shm[q].buf)
            self.img[q] = view.ctypes.data
        if self.dist is not None:
            # groups must be created by every rank, in the same order (the step list is the same on all ranks)
            for c in self.colls:
                key = (c["lo"], c["hi"])
                if key not in self.groups:
                    self.groups[key] = None if key == (0, self.world) else self.dist.new_group(ranks=list(range(*key)))
        self.connected = True

    def _meet(self, c):
        if self.dist is not None and self.world > 1:
            self.dist.barrier(group=self.groups[(c["lo"], c["hi"])])

    def factorize(self):
        assert self.connected
        self.X[:] = 0.0
        self.bad = 0
        self.pushed_bytes = 0
        self.L.sfcheck_begin(self.h, self.X.ctypes.data)
        for k, sg in enumerate(self.segments):
            for i in range(sg["coll_first"], sg["coll_first"] + sg["coll_count"]):
                c = self.colls[i]
                if not (c["lo"] <= self.rank < c["hi"]):
                    continue
                if c["type"] == 0:                       # barrier: the group's remote adds are in place
                    self._meet(c)
                else:                                    # push: root copies the block into the group's images, then flags
                    if c["root"] == self.rank:
                        assert self.L.sfcheck_push(self.h, i, self.img) == 0
                        self.pushed_bytes += 8 * c["rows"] * c["cols"] * (c["hi"] - c["lo"] - 1)
                    self._meet(c)
            self.bad += self.L.sfcheck_run_segment(self.h, k, self.img)
        # the device side needs no closing step (every remote write is followed by a flag its target waits for); the
        # processes of the test must not unmap their segments while a slower peer still works in them
        if self.dist is not None and self.world > 1:
            self.dist.barrier()

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
        self.X = None
        for q, m in enumerate(self.peers_shm):
            if m is not None:
                m.close()
        self.peers_shm = []
        if self.shm is not None:
            self.shm.close()
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
            self.shm = None
