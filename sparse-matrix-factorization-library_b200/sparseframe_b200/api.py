"""Reference-shaped host API: compress / analyze / factorize / solve_supernodal / validate."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi

# devSlotSize the reference computes for one B200 (GPU_SPLIT_LIMIT=4 virtual GPUs x 8 slots, CH.c:81-87,199);
# the symbolic structure depends on it (CH.c:1476-1496, 1556-1574), so it is pinned and recorded with every number.
DEFAULT_DEVSLOT = 5_380_000_000

CHOLESKY, LU = 0, 1


def _kind(kind) -> int:
    if kind in (LU, "lu", "LU"):
        return LU
    if kind in (CHOLESKY, "cholesky", "chol", "Cholesky"):
        return CHOLESKY
    raise ValueError("kind must be 'cholesky' or 'lu'")


@dataclass
class Matrix:
    """Compressed-column matrix as SparseFrame_compress leaves it (CH.c:526-587)."""
    n: int
    Cp: np.ndarray
    Ci: np.ndarray
    Cx: np.ndarray
    symmetric: bool


def compress(n, rows, cols, vals, symmetric=True) -> Matrix:
    """Triplets -> CSC by a stable counting sort on the column (CH.c:560-576): file order is kept inside each
    column, nothing is summed or sorted.  Complex values make a complex matrix (matrix_info->isComplex)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.complex128 if np.iscomplexobj(vals) else np.float64)
    order = np.argsort(cols, kind="stable")
    cp = np.zeros(n + 1, dtype=np.int64)
    np.add.at(cp, cols + 1, 1)
    return Matrix(int(n), np.cumsum(cp), np.ascontiguousarray(rows[order]), np.ascontiguousarray(vals[order]),
                  bool(symmetric))


class Symbolic:
    """Result of symbolic analysis (SparseFrame_analyze, CH.c:1916-1978 / LU.c:2233-2458)."""

    _INT64 = ("Order", "Perm", "Parent", "Post", "ColCount", "Lp", "Li", "LTp", "LTi", "Up", "Ui", "UTp", "UTi", "Super",
              "SuperMap", "Sparent", "Lsip", "Lsxp", "Lsi", "ST_Map", "ST_Pointer", "ST_Index", "LeafQueue",
              "Aoffset", "Moffset", "CPCTp", "CPCTi")
    _F64 = ("Lx", "LTx", "Ux", "UTx")

    def __init__(self, handle, kind, devslot):
        self._h = handle
        self.kind = kind
        self.devslot = devslot
        self.complex = bool(_capi.lib().sfb200_symbolic_int(handle, b"complex") == 1)

    def int(self, name) -> int:
        return int(_capi.lib().sfb200_symbolic_int(self._h, name.encode()))

    def array(self, name) -> np.ndarray:
        cnt = C.c_int64()
        p = _capi.lib().sfb200_symbolic_array(self._h, name.encode(), C.byref(cnt))
        dt = np.float64 if name in self._F64 else np.int64
        if not p or cnt.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (cnt.value * 8)).from_address(p)
        a = np.frombuffer(buf, dtype=dt, count=cnt.value).copy()
        return a.view(np.complex128) if (self.complex and name in self._F64) else a

    def double(self, name) -> float:
        return float(_capi.lib().sfb200_symbolic_double(self._h, name.encode()))

    @property
    def flops(self) -> float:
        return float(_capi.lib().sfb200_symbolic_flops(self._h))

    def close(self):
        if self._h:
            _capi.lib().sfb200_symbolic_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def analyze(A: Matrix, kind="cholesky", devslot=DEFAULT_DEVSLOT, order=None) -> Symbolic:
    """Symbolic analysis; `order` (the "Order" array of an earlier analysis of the same pattern, or any permutation)
    replaces the METIS call."""
    L = _capi.lib()
    h = C.c_void_p()
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.int64)
        if order.shape != (A.n,):
            raise ValueError("order must have n entries")
    fn = L.sfb200_analyze_complex if np.iscomplexobj(A.Cx) else L.sfb200_analyze_ordered
    rc = fn(_kind(kind), int(A.symmetric), A.n, A.Cp.ctypes.data, A.Ci.ctypes.data,
            A.Cx.ctypes.data, int(devslot), order.ctypes.data if order is not None else None, C.byref(h))
    if rc != 0:
        raise RuntimeError("sfb200_analyze failed (%d): %s" % (rc, _capi.last_error()))
    return Symbolic(h, _kind(kind), int(devslot))


def wrap_structure(kind, n, Super, Lsip, Lsxp, Lsi, Lp, Li, Up=None, Ui=None, is_complex=False) -> Symbolic:
    """Symbolic object around a structure computed elsewhere (e.g. by the reference's own analysis)."""
    L = _capi.lib()
    arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (Super, Lsip, Lsxp, Lsi, Lp, Li)]
    up = np.ascontiguousarray(Up, dtype=np.int64) if Up is not None else None
    ui = np.ascontiguousarray(Ui, dtype=np.int64) if Ui is not None else None
    h = C.c_void_p()
    wrap = L.sfb200_symbolic_wrap_complex if is_complex else L.sfb200_symbolic_wrap
    rc = wrap(_kind(kind), int(n), len(arrs[0]) - 1, *[a.ctypes.data for a in arrs],
                up.ctypes.data if up is not None else None,
                ui.ctypes.data if ui is not None else None, C.byref(h))
    if rc != 0:
        raise RuntimeError("sfb200_symbolic_wrap failed (%d): %s" % (rc, _capi.last_error()))
    return Symbolic(h, _kind(kind), 0)


def device_count() -> int:
    return int(_capi.lib().sfb200_device_count())


def fp64_peak(device=0, which="dmma", iters=4096) -> float:
    code = {"dfma": 0, "dmma": 1, "dmma_smem_tile": 2, "dmma_reg_tile": 3}[which]
    return float(_capi.lib().sfb200_probe_fp64_peak(int(device), code, int(iters)))


class Plan:
    """Device-resident numeric schedule for one symbolic structure."""

    def __init__(self, sym: Symbolic, device=0, rank=0, world=1, _handle=None):
        L = _capi.lib()
        self._owned = _handle is None
        if _handle is None:
            h = C.c_void_p()
            rc = L.sfb200_plan_create_part(sym._h, int(device), int(rank), int(world), C.byref(h))
            if rc != 0:
                raise RuntimeError("sfb200_plan_create failed (%d): %s" % (rc, _capi.last_error()))
        else:
            h = C.c_void_p(_handle)           # a rank of a Group: freed with the group
        self._h = h
        self.sym = sym
        self.rank, self.world = int(rank), int(world)

    # ---- etree-partitioned execution (one rank of `world`): see parallel.py ------------------------------------
    def double(self, name) -> float:
        return float(_capi.lib().sfb200_plan_double(self._h, name.encode()))

    def segment_info(self, k):
        out = np.zeros(5, dtype=np.int64)
        _capi.lib().sfb200_plan_segment_info(self._h, int(k), out.ctypes.data)
        return dict(coll_first=int(out[0]), coll_count=int(out[1]), launch_first=int(out[2]), launch_count=int(out[3]),
                    level=int(out[4]))

    def coll_info(self, i):
        out = np.zeros(8, dtype=np.int64)
        _capi.lib().sfb200_plan_coll_info(self._h, int(i), out.ctypes.data)
        return dict(zip(("type", "root", "lo", "hi", "offset", "rows", "cols", "ld"), (int(v) for v in out)))

    def owner(self) -> np.ndarray:
        out = np.zeros(self.int("nsuper"), dtype=np.int32)
        _capi.lib().sfb200_plan_owner(self._h, out.ctypes.data)
        return out

    def panel_offsets(self) -> np.ndarray:
        """Offsets of the panels in this rank's image (-1: not stored here); last entry = elements of the image."""
        out = np.zeros(self.int("nsuper") + 1, dtype=np.int64)
        _capi.lib().sfb200_plan_panel_offsets(self._h, out.ctypes.data)
        return out

    def factor_ptr(self) -> int:
        return int(_capi.lib().sfb200_plan_factor_ptr(self._h))

    def exchange_handle(self) -> bytes:
        """128 bytes another process needs to map this rank's image and flags (CUDA IPC)."""
        buf = C.create_string_buffer(128)
        rc = _capi.lib().sfb200_plan_exchange_handle(self._h, buf)
        if rc != 0:
            raise RuntimeError("sfb200_plan_exchange_handle failed (%d): %s" % (rc, _capi.last_error()))
        return buf.raw

    def connect(self, handles: bytes):
        """handles = the exchange handles of all ranks, concatenated in rank order."""
        if len(handles) != 128 * self.world:
            raise ValueError("need %d bytes of handles" % (128 * self.world))
        rc = _capi.lib().sfb200_plan_connect(self._h, C.c_char_p(handles))
        if rc != 0:
            raise RuntimeError("sfb200_plan_connect failed (%d): %s" % (rc, _capi.last_error()))

    def adopt_factor(self, d_lx_ptr: int, d_ux_ptr: int = 0):
        rc = _capi.lib().sfb200_plan_adopt_factor(self._h, d_lx_ptr, d_ux_ptr or None)
        if rc != 0:
            raise RuntimeError("sfb200_plan_adopt_factor failed (%d): %s" % (rc, _capi.last_error()))

    def int(self, name) -> int:
        return int(_capi.lib().sfb200_plan_int(self._h, name.encode()))

    def factorize(self, Lx, Ux=None, out=None, download=True):
        """Host-buffer factorization (what SparseFrame_factorize maps to).  Returns Lsx in the reference layout."""
        L = _capi.lib()
        Lx = np.ascontiguousarray(Lx, dtype=np.float64)
        ux_p = None
        if Ux is not None:
            Ux = np.ascontiguousarray(Ux, dtype=np.float64)
            ux_p = Ux.ctypes.data
        if download and out is None:
            out = np.empty(self.int("xsize"), dtype=np.float64)
        rc = L.sfb200_factorize(self._h, Lx.ctypes.data, ux_p, out.ctypes.data if download else None)
        if rc != 0:
            raise RuntimeError("sfb200_factorize failed (%d): %s" % (rc, _capi.last_error()))
        return out

    def factorize_device(self, d_lx_ptr: int, d_ux_ptr: int = 0):
        rc = _capi.lib().sfb200_factorize_device(self._h, d_lx_ptr, d_ux_ptr or None)
        if rc != 0:
            raise RuntimeError("sfb200_factorize_device failed (%d): %s" % (rc, _capi.last_error()))

    def sync(self):
        rc = _capi.lib().sfb200_plan_sync(self._h)
        if rc != 0:
            raise RuntimeError("factorization failed (%d): %s" % (rc, _capi.last_error()))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.int("xsize"), dtype=np.float64)
        rc = _capi.lib().sfb200_download_factor(self._h, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("sfb200_download_factor failed (%d): %s" % (rc, _capi.last_error()))
        return out

    def set_stream(self, cuda_stream_ptr: int):
        rc = _capi.lib().sfb200_plan_set_stream(self._h, cuda_stream_ptr)
        if rc != 0:
            raise RuntimeError("sfb200_plan_set_stream failed (%d): %s" % (rc, _capi.last_error()))

    KERNELS = {0: "k_small<64>", 1: "k_small<256>", 2: "k_diag<potrf>", 3: "k_diag<getrf>", 4: "k_trsm",
               5: "k_gemm<128x128>", 6: "k_gemm<64x64>", 7: "k_gemm_ws<128x128>", 8: "k_trsm_mma"}

    def profile(self):
        """Instrumented replay: {kernel: dict(ms, launches, flops, bytes)} for the last factorization's values."""
        L = _capi.lib()
        rc = L.sfb200_plan_profile(self._h)
        if rc != 0:
            raise RuntimeError("sfb200_plan_profile failed (%d): %s" % (rc, _capi.last_error()))
        out = {}
        for k, name in self.KERNELS.items():
            n = L.sfb200_plan_profile_get(self._h, k, 1)
            if n > 0:
                out[name] = dict(ms=L.sfb200_plan_profile_get(self._h, k, 0), launches=int(n),
                                 flops=L.sfb200_plan_profile_get(self._h, k, 2),
                                 bytes=L.sfb200_plan_profile_get(self._h, k, 3))
        return out

    def profile_launches(self):
        """Per-launch rows of the last profile(): kind, level, grid, ops, ms, flops, kmax, kmin."""
        rows = []
        out = np.zeros(8)
        for i in range(self.int("nlaunches")):
            if _capi.lib().sfb200_plan_profile_launch(self._h, i, out.ctypes.data) != 0:
                break
            rows.append(out.copy())
        return np.array(rows)

    def solve(self, b):
        """x = A^-1 b on the GPU with the resident factor (permuted ordering); b: (n,) or (n, nrhs) Fortran-ordered."""
        x = np.array(b, dtype=np.float64, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        rc = _capi.lib().sfb200_solve(self._h, x.ctypes.data, nrhs)
        if rc != 0:
            raise RuntimeError("sfb200_solve failed (%d): %s" % (rc, _capi.last_error()))
        return x

    def residual(self, x, b):
        """dict(anorm, bnorm, xnorm, rnorm, residual) of the permuted system, computed on the GPU."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        out = np.zeros(5)
        rc = _capi.lib().sfb200_residual(self._h, x.ctypes.data, b.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("sfb200_residual failed (%d): %s" % (rc, _capi.last_error()))
        return dict(zip(("anorm", "bnorm", "xnorm", "rnorm", "residual"), out.tolist()))

    def last_ms(self, which=0) -> float:
        return float(_capi.lib().sfb200_last_ms(self._h, which))

    def last_launches(self) -> int:
        return int(_capi.lib().sfb200_last_launches(self._h))

    def close(self):
        if self._h:
            if self._owned:
                _capi.lib().sfb200_plan_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZPlan:
    """Numeric plan of a COMPLEX matrix (sfb200_zplan_*): the FP64 engine on the real embedding of the matrix, complex
    panels read back in the reference's layout (csrc/complex_embed.hpp).  Values are complex128 arrays."""

    def __init__(self, sym: Symbolic, device=0, row_phase=None):
        L = _capi.lib()
        h = C.c_void_p()
        rc = L.sfb200_zplan_create(sym._h, int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("sfb200_zplan_create failed (%d): %s" % (rc, _capi.last_error()))
        self._h = h
        self.sym = sym
        if row_phase is not None:
            L.sfb200_zplan_set_row_phase(h, int(bool(row_phase)))
        # the plan of the embedded real structure, for inspection (freed with the zplan)
        self.real = Plan(sym, device, _handle=L.sfb200_zplan_real(h))

    def int(self, name) -> int:
        return int(_capi.lib().sfb200_zplan_int(self._h, name.encode()))

    def factorize(self, Lz, Uz=None, download=True):
        L = _capi.lib()
        Lz = np.ascontiguousarray(Lz, dtype=np.complex128)
        uz_p = None
        if Uz is not None:
            Uz = np.ascontiguousarray(Uz, dtype=np.complex128)
            uz_p = Uz.ctypes.data
        out = np.empty(self.int("xsize"), dtype=np.complex128) if download else None
        rc = L.sfb200_zfactorize(self._h, Lz.ctypes.data, uz_p, out.ctypes.data if download else None)
        if rc != 0:
            raise RuntimeError("sfb200_zfactorize failed (%d): %s" % (rc, _capi.last_error()))
        return out

    def download(self):
        out = np.empty(self.int("xsize"), dtype=np.complex128)
        rc = _capi.lib().sfb200_zdownload_factor(self._h, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("sfb200_zdownload_factor failed (%d): %s" % (rc, _capi.last_error()))
        return out

    def solve(self, b):
        x = np.array(b, dtype=np.complex128, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        rc = _capi.lib().sfb200_zsolve(self._h, x.ctypes.data, nrhs)
        if rc != 0:
            raise RuntimeError("sfb200_zsolve failed (%d): %s" % (rc, _capi.last_error()))
        return x

    def residual(self, x, b):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        b = np.ascontiguousarray(b, dtype=np.complex128)
        out = np.zeros(5)
        rc = _capi.lib().sfb200_zresidual(self._h, x.ctypes.data, b.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("sfb200_zresidual failed (%d): %s" % (rc, _capi.last_error()))
        return dict(zip(("anorm", "bnorm", "xnorm", "rnorm", "residual"), out.tolist()))

    def last_ms(self) -> float:
        return float(_capi.lib().sfb200_zlast_ms(self._h))

    def close(self):
        if self._h:
            self.real._h = None
            _capi.lib().sfb200_zplan_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZGroup:
    """A COMPLEX matrix on several GPUs of one box (sfb200_zgroup_*): the etree partition of the embedded structure as an
    in-process group, values expanded and complex panels read back on the host."""

    def __init__(self, sym: Symbolic, devices):
        L = _capi.lib()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = L.sfb200_zgroup_create(sym._h, len(devices), devs, C.byref(h))
        if rc != 0:
            raise RuntimeError("sfb200_zgroup_create failed (%d): %s" % (rc, _capi.last_error()))
        self._h = h
        self.sym = sym
        self.world = len(devices)

    def factorize(self, Lz, Uz=None, download=True):
        L = _capi.lib()
        Lz = np.ascontiguousarray(Lz, dtype=np.complex128)
        uz_p = None
        if Uz is not None:
            Uz = np.ascontiguousarray(Uz, dtype=np.complex128)
            uz_p = Uz.ctypes.data
        out = np.full(self.sym.int("xsize"), np.nan + 0j, dtype=np.complex128) if download else None
        rc = L.sfb200_zgroup_factorize(self._h, Lz.ctypes.data, uz_p, out.ctypes.data if download else None)
        if rc != 0:
            raise RuntimeError("sfb200_zgroup_factorize failed (%d): %s" % (rc, _capi.last_error()))
        return out

    def last_ms(self, which=0) -> float:
        return float(_capi.lib().sfb200_zgroup_last_ms(self._h, which))

    def close(self):
        if self._h:
            _capi.lib().sfb200_zgroup_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """All ranks of an etree partition inside one process, rank r on devices[r] (sfb200_group_*): what the drop-in's
    SparseFrame_factorize drives when numGPU > 1 (the reference: one host thread per GPU, CH.c:2267-2290)."""

    def __init__(self, sym: Symbolic, devices):
        L = _capi.lib()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = L.sfb200_group_create(sym._h, len(devices), devs, C.byref(h))
        if rc != 0:
            raise RuntimeError("sfb200_group_create failed (%d): %s" % (rc, _capi.last_error()))
        self._h = h
        self.sym = sym
        self.world = len(devices)
        self.plans = [Plan(sym, devices[r], r, self.world, _handle=L.sfb200_group_plan(h, r)) for r in range(self.world)]

    def factorize(self, Lx, Ux=None, out=None, download=True):
        """Host buffers in, factor in the reference layout out (every rank writes the panels it owns)."""
        L = _capi.lib()
        Lx = np.ascontiguousarray(Lx, dtype=np.float64)
        ux_p = None
        if Ux is not None:
            Ux = np.ascontiguousarray(Ux, dtype=np.float64)
            ux_p = Ux.ctypes.data
        if download and out is None:
            out = np.full(self.plans[0].int("xsize"), np.nan)
        rc = L.sfb200_group_factorize(self._h, Lx.ctypes.data, ux_p, out.ctypes.data if download else None)
        if rc != 0:
            raise RuntimeError("sfb200_group_factorize failed (%d): %s" % (rc, _capi.last_error()))
        return out

    def last_ms(self, which=0) -> float:
        return float(_capi.lib().sfb200_group_last_ms(self._h, which))

    def close(self):
        if self._h:
            for p in self.plans:
                p.close()
            _capi.lib().sfb200_group_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def factorize(sym: Symbolic, device=0):
    """analyze -> factorize in one call, values taken from the analysis (Lx / Ux)."""
    plan = Plan(sym, device)
    ux = sym.array("Ux") if (sym.kind == LU and sym.int("symmetric") == 0) else None
    lsx = plan.factorize(sym.array("Lx"), ux)
    return plan, lsx


# ---------------------------------------------------------------------------------------------------------------
# consumers of the factor (host, numpy): SparseFrame_solve_supernodal / SparseFrame_validate
# ---------------------------------------------------------------------------------------------------------------
def solve_supernodal(kind, Super, Lsip, Lsxp, Lsi, Lsx, b):
    """Forward/backward substitution over the supernodal factor (CH.c:3036-3139; LU.c:3592-3700), blocked per
    supernode instead of scalar."""
    from scipy.linalg import solve_triangular
    lu = _kind(kind) == LU
    x = np.array(b, dtype=np.float64, copy=True)
    ns = len(Super) - 1
    for s in range(ns):
        nc = Super[s + 1] - Super[s]
        nr = Lsip[s + 1] - Lsip[s]
        ld = 2 * nr - nc if lu else nr
        P = Lsx[Lsxp[s]:Lsxp[s] + nc * ld].reshape(nc, ld).T          # ld x nc panel
        rows = Lsi[Lsip[s]:Lsip[s + 1]]
        xs = solve_triangular(P[:nc, :nc], x[rows[:nc]], lower=True, unit_diagonal=lu, check_finite=False)
        x[rows[:nc]] = xs
        if nr > nc:
            x[rows[nc:]] -= P[nc:nr, :] @ xs
    for s in range(ns - 1, -1, -1):
        nc = Super[s + 1] - Super[s]
        nr = Lsip[s + 1] - Lsip[s]
        ld = 2 * nr - nc if lu else nr
        P = Lsx[Lsxp[s]:Lsxp[s] + nc * ld].reshape(nc, ld).T
        rows = Lsi[Lsip[s]:Lsip[s + 1]]
        rhs = x[rows[:nc]]
        if nr > nc:
            off = P[nr:, :] if lu else P[nc:nr, :]                     # LU: U12^T block, Cholesky: L21
            rhs = rhs - off.T @ x[rows[nc:]]
        if lu:
            x[rows[:nc]] = solve_triangular(P[:nc, :nc], rhs, lower=False, check_finite=False)
        else:
            x[rows[:nc]] = solve_triangular(P[:nc, :nc], rhs, lower=True, trans="T", check_finite=False)
    return x


def validate(kind, n, Lp, Li, Lx, Super, Lsip, Lsxp, Lsi, Lsx, Up=None, Ui=None, Ux=None):
    """The reference's self-check (CH.c:3141-3266): b_i = 1 + i/n, solve, r = A x - b in the PERMUTED system,
    returns (residual as the reference prints it, relative backward error ||Ax-b|| / (||A|| ||x||))."""
    import scipy.sparse as sp
    b = 1.0 + np.arange(n, dtype=np.float64) / n
    x = solve_supernodal(kind, Super, Lsip, Lsxp, Lsi, Lsx, b)
    cols = np.repeat(np.arange(n), np.diff(Lp))
    lower = sp.csc_matrix((Lx, Li, Lp), shape=(n, n))
    if Up is None:
        offd = sp.coo_matrix((Lx, (cols, Li)), shape=(n, n)).tocsc()
        offd.setdiag(0)
        A = lower + offd
    else:
        urow = np.repeat(np.arange(n), np.diff(Up))
        strict = Li > cols
        A = sp.coo_matrix((Lx[strict], (Li[strict], cols[strict])), shape=(n, n)).tocsc() + \
            sp.coo_matrix((Ux, (urow, Ui)), shape=(n, n)).tocsc()
    r = A @ x - b
    anorm1 = abs(A).sum(axis=0).max()
    xinf, binf, rinf = np.abs(x).max(), np.abs(b).max(), np.abs(r).max()
    residual = rinf / (anorm1 * xinf + binf)
    backward = rinf / (anorm1 * xinf)
    return float(residual), float(backward)


def lower_trapezoid_mask(kind, Super, Lsip, Lsxp, xsize):
    """Entries of Lsx the reference defines.  Cholesky: the lower trapezoid of every panel (the strict upper
    triangle of the diagonal block is never written, CH.c:2135).  LU: everything."""
    mask = np.ones(int(xsize), dtype=bool)
    if _kind(kind) == LU:
        return mask
    for s in range(len(Super) - 1):
        nc = int(Super[s + 1] - Super[s])
        nr = int(Lsip[s + 1] - Lsip[s])
        if nc > 1:
            blk = mask[Lsxp[s]:Lsxp[s] + nc * nr].reshape(nc, nr)
            blk[:, :nc] = np.triu(np.ones((nc, nc), dtype=bool))
    return mask


def entrywise_rel_diff(a, b):
    """max |a-b| / max(|a|,|b|) with 0/0 := 0 (explicit zeros of relaxed amalgamation must stay exactly zero)."""
    den = np.maximum(np.abs(a), np.abs(b))
    d = np.abs(a - b)
    nz = den > 0
    if not nz.any():
        return 0.0
    return float((d[nz] / den[nz]).max())
