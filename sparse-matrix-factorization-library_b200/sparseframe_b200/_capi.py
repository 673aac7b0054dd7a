"""ctypes binding of lib/libsfb200.so (C ABI: include/sfb200.h).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# SFB200_LIBRARY: another build of the same library (kernel experiments: scripts/perf_sweep.py); never a fallback
LIB_PATH = os.environ.get("SFB200_LIBRARY") or os.path.join(os.path.dirname(_PKG), "lib", "libsfb200.so")

_lib = None

SYMBOLS = [
    "sfb200_version", "sfb200_last_error", "sfb200_analyze", "sfb200_symbolic_wrap", "sfb200_symbolic_int",
    "sfb200_symbolic_array", "sfb200_symbolic_flops", "sfb200_symbolic_free", "sfb200_device_count",
    "sfb200_plan_create", "sfb200_plan_free", "sfb200_plan_int", "sfb200_factorize", "sfb200_factorize_device",
    "sfb200_plan_sync", "sfb200_download_factor", "sfb200_last_ms", "sfb200_last_launches", "sfb200_host_alloc",
    "sfb200_host_free", "sfb200_probe_fp64_peak", "sfb200_plan_set_stream", "sfb200_plan_profile",
    "sfb200_plan_profile_get", "sfb200_plan_create_part", "sfb200_plan_double", "sfb200_plan_segment_info",
    "sfb200_plan_coll_info", "sfb200_plan_owner", "sfb200_plan_factor_ptr", "sfb200_plan_exchange_handle",
    "sfb200_plan_connect", "sfb200_group_create", "sfb200_group_size", "sfb200_group_plan", "sfb200_group_factorize",
    "sfb200_group_last_ms", "sfb200_group_free", "sfb200_plan_adopt_factor", "sfb200_plan_panel_offsets", "sfb200_factorize_owned", "sfb200_download_owned",
    "sfb200_plan_owned_runs", "sfb200_plan_profile_launch", "sfb200_solve", "sfb200_solve_device", "sfb200_residual",
    "sfb200_analyze_ordered", "sfb200_symbolic_double", "sfb200_set_ordering_processes",
    # complex matrices (real embedding, csrc/zengine.cu)
    "sfb200_analyze_complex", "sfb200_symbolic_wrap_complex", "sfb200_zplan_create", "sfb200_zplan_free", "sfb200_zplan_real",
    "sfb200_zplan_int", "sfb200_zplan_set_row_phase", "sfb200_zfactorize", "sfb200_zfactorize_device", "sfb200_zplan_sync",
    "sfb200_zdownload_factor", "sfb200_zsolve", "sfb200_zresidual", "sfb200_zlast_ms",
    "sfb200_zgroup_create", "sfb200_zgroup_real", "sfb200_zgroup_factorize", "sfb200_zgroup_last_ms", "sfb200_zgroup_free",
]


class LibraryMissing(RuntimeError):
    pass


def lib():
    """Load the native library.  There is no Python/CPU fallback: a missing extension is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "native library %s is missing: run `make -C sparse-matrix-factorization-library_b200` "
            "(or __graft_entry__.build()); this engine has no fallback path" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, u64, dbl, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_double, C.c_int
    L.sfb200_version.restype = C.c_char_p
    L.sfb200_last_error.restype = C.c_char_p
    L.sfb200_analyze.argtypes = [ci, ci, i64, vp, vp, vp, u64, C.POINTER(vp)]
    L.sfb200_analyze_ordered.argtypes = [ci, ci, i64, vp, vp, vp, u64, vp, C.POINTER(vp)]
    L.sfb200_symbolic_double.restype = dbl
    L.sfb200_symbolic_double.argtypes = [vp, C.c_char_p]
    L.sfb200_symbolic_wrap.argtypes = [ci, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.sfb200_symbolic_int.restype = i64
    L.sfb200_symbolic_int.argtypes = [vp, C.c_char_p]
    L.sfb200_symbolic_array.restype = vp
    L.sfb200_symbolic_array.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
    L.sfb200_symbolic_flops.restype = dbl
    L.sfb200_symbolic_flops.argtypes = [vp]
    L.sfb200_symbolic_free.argtypes = [vp]
    L.sfb200_symbolic_free.restype = None
    L.sfb200_device_count.restype = ci
    L.sfb200_plan_create.argtypes = [vp, ci, C.POINTER(vp)]
    L.sfb200_plan_free.argtypes = [vp]
    L.sfb200_plan_free.restype = None
    L.sfb200_plan_int.restype = i64
    L.sfb200_plan_int.argtypes = [vp, C.c_char_p]
    L.sfb200_factorize.argtypes = [vp, vp, vp, vp]
    L.sfb200_factorize_device.argtypes = [vp, vp, vp]
    L.sfb200_plan_sync.argtypes = [vp]
    L.sfb200_download_factor.argtypes = [vp, vp]
    L.sfb200_last_ms.restype = dbl
    L.sfb200_last_ms.argtypes = [vp, ci]
    L.sfb200_last_launches.restype = i64
    L.sfb200_last_launches.argtypes = [vp]
    L.sfb200_host_alloc.restype = vp
    L.sfb200_host_alloc.argtypes = [C.c_size_t]
    L.sfb200_host_free.argtypes = [vp]
    L.sfb200_host_free.restype = None
    L.sfb200_probe_fp64_peak.restype = dbl
    L.sfb200_probe_fp64_peak.argtypes = [ci, ci, ci]
    L.sfb200_plan_create_part.argtypes = [vp, ci, ci, ci, C.POINTER(vp)]
    L.sfb200_plan_double.restype = dbl
    L.sfb200_plan_double.argtypes = [vp, C.c_char_p]
    L.sfb200_plan_segment_info.argtypes = [vp, ci, vp]
    L.sfb200_plan_coll_info.argtypes = [vp, ci, vp]
    L.sfb200_plan_owner.argtypes = [vp, vp]
    L.sfb200_plan_factor_ptr.restype = vp
    L.sfb200_plan_factor_ptr.argtypes = [vp]
    L.sfb200_plan_panel_offsets.argtypes = [vp, vp]
    L.sfb200_plan_profile_launch.argtypes = [vp, ci, vp]
    L.sfb200_plan_exchange_handle.argtypes = [vp, vp]
    L.sfb200_plan_connect.argtypes = [vp, vp]
    L.sfb200_group_create.argtypes = [vp, ci, vp, C.POINTER(vp)]
    L.sfb200_group_size.argtypes = [vp]
    L.sfb200_group_plan.restype = vp
    L.sfb200_group_plan.argtypes = [vp, ci]
    L.sfb200_group_factorize.argtypes = [vp, vp, vp, vp]
    L.sfb200_group_last_ms.restype = dbl
    L.sfb200_group_last_ms.argtypes = [vp, ci]
    L.sfb200_group_free.argtypes = [vp]
    L.sfb200_group_free.restype = None
    L.sfb200_plan_adopt_factor.argtypes = [vp, vp, vp]
    L.sfb200_factorize_owned.argtypes = [vp, vp, vp, vp]
    L.sfb200_download_owned.argtypes = [vp, vp]
    L.sfb200_plan_owned_runs.restype = i64
    L.sfb200_plan_owned_runs.argtypes = [vp, vp]
    L.sfb200_solve.argtypes = [vp, vp, ci]
    L.sfb200_solve_device.argtypes = [vp, vp, ci]
    L.sfb200_residual.argtypes = [vp, vp, vp, vp]
    L.sfb200_plan_set_stream.argtypes = [vp, vp]
    L.sfb200_plan_profile.argtypes = [vp]
    L.sfb200_plan_profile_get.restype = dbl
    L.sfb200_plan_profile_get.argtypes = [vp, ci, ci]
    L.sfb200_set_ordering_processes.argtypes = [ci]
    L.sfb200_analyze_complex.argtypes = [ci, ci, i64, vp, vp, vp, u64, vp, C.POINTER(vp)]
    L.sfb200_symbolic_wrap_complex.argtypes = L.sfb200_symbolic_wrap.argtypes
    L.sfb200_zplan_create.argtypes = [vp, ci, C.POINTER(vp)]
    L.sfb200_zplan_free.argtypes = [vp]
    L.sfb200_zplan_free.restype = None
    L.sfb200_zplan_real.restype = vp
    L.sfb200_zplan_real.argtypes = [vp]
    L.sfb200_zplan_int.restype = i64
    L.sfb200_zplan_int.argtypes = [vp, C.c_char_p]
    L.sfb200_zplan_set_row_phase.argtypes = [vp, ci]
    L.sfb200_zfactorize.argtypes = [vp, vp, vp, vp]
    L.sfb200_zfactorize_device.argtypes = [vp, vp, vp]
    L.sfb200_zplan_sync.argtypes = [vp]
    L.sfb200_zdownload_factor.argtypes = [vp, vp]
    L.sfb200_zsolve.argtypes = [vp, vp, ci]
    L.sfb200_zresidual.argtypes = [vp, vp, vp, vp]
    L.sfb200_zlast_ms.restype = dbl
    L.sfb200_zlast_ms.argtypes = [vp]
    L.sfb200_zgroup_create.argtypes = [vp, ci, vp, C.POINTER(vp)]
    L.sfb200_zgroup_real.restype = vp
    L.sfb200_zgroup_real.argtypes = [vp]
    L.sfb200_zgroup_factorize.argtypes = [vp, vp, vp, vp]
    L.sfb200_zgroup_last_ms.restype = dbl
    L.sfb200_zgroup_last_ms.argtypes = [vp, ci]
    L.sfb200_zgroup_free.argtypes = [vp]
    L.sfb200_zgroup_free.restype = None
    _lib = L
    return L


def last_error() -> str:
    return lib().sfb200_last_error().decode()
