"""ctypes binding of include/cannoles_b200.h.

``load()`` opens the product library ``csrc/libcannoles_b200.so`` (nvcc build for sm_100a) and
nothing else: if it is missing the import fails loudly -- there is no CPU fallback.
``bind_library(path)`` only attaches the prototypes to an already chosen shared object; the
CPU-emulator build used by ``tests/hostsim`` goes through it explicitly from test code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcannoles_b200.so")

p64 = C.POINTER(C.c_int64)
p32 = C.POINTER(C.c_int32)
pd = C.POINTER(C.c_double)
pu8 = C.POINTER(C.c_uint8)


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "N", "nnz", "nnzA", "nnzL", "nnzL_store", "cb_store", "nsuper", "nlevels", "max_front",
        "max_width", "n_small", "n_large", "launches_factor", "launches_solve")] + [
        (n, C.c_double) for n in ("flops", "flops_store", "t_order", "t_symbolic", "t_plan",
                                  "bytes_device")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class NLSParams(C.Structure):
    """b2_nls_params_t"""
    _fields_ = [(n, C.c_double) for n in (
        "eig_tol", "delta_min", "kappa_dec", "kappa_inc", "kappa_largeinc", "rho0", "rho_max", "rho_min",
        "gamma_A", "atol", "rtol", "Fatol", "Frtol", "delta_dec", "cgls_tol")] + [
        (n, C.c_int32) for n in ("max_iter", "max_eval", "max_inner", "always_accept_extrapolation",
                                 "use_initial_multiplier", "reserved")]


class DenseNLS(C.Structure):
    """b2_dense_nls_t"""
    _fields_ = [(n, C.c_int64) for n in ("n", "m", "ncon", "shared_model")] + [
        (n, pd) for n in ("At", "Bt", "Ct", "y", "e", "x0", "y0")]


# every symbol include/cannoles_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "b2_last_error", "b2_version", "b2_device_count", "b2_analyze", "b2_factorize",
    "b2_refactorize_shift", "b2_factorize_retry", "b2_solve", "b2_factorize_dev", "b2_solve_dev", "b2_register_host",
    "b2_unregister_host", "b2_stats", "b2_last_timings", "b2_timer_start", "b2_timer_stop", "b2_last_sweeps", "b2_profile", "b2_front_sizes", "b2_get_perm", "b2_get_csc",
    "b2_get_nzval", "b2_get_d", "b2_set_option", "b2_free",
    "b2b_analyze", "b2b_factorize", "b2b_refactorize_shift", "b2b_solve", "b2b_factorize_dev",
    "b2b_solve_dev", "b2b_factor_solve", "b2b_factor_solve_dev", "b2b_stats", "b2b_last_ms",
    "b2b_timer_start", "b2b_timer_stop", "b2b_get_perm", "b2b_get_d", "b2b_free",
    "b2_nls_default_params", "b2b_nls_record_len", "b2b_nls_dense_solve_dev", "b2b_nls_dense_solve",
    "b2b_nls_dense_submit", "b2b_nls_wait",
    "b2_dev_malloc", "b2_dev_free", "b2_dev_upload", "b2_dev_download", "b2_dev_sync",
    "b2_host_register", "b2_host_unregister",
    "b2_measure_dgemm", "b2_measure_hbm",
]


def bind_library(path: str):
    lib = C.CDLL(path)
    vp = C.c_void_p
    pi = C.POINTER(C.c_int)
    lib.b2_last_error.restype = C.c_char_p
    lib.b2_analyze.argtypes = [C.c_int64, C.c_int64, p64, p64, C.c_int64, C.c_int64, C.c_int64,
                               C.c_int, p64, C.c_int, C.POINTER(vp)]
    lib.b2_factorize.argtypes = [vp, pd, C.c_double, p64, p64, p64, pi]
    lib.b2_factorize_dev.argtypes = [vp, vp, C.c_double, p64, p64, p64, pi]
    lib.b2_refactorize_shift.argtypes = [vp, C.c_double, C.c_double, C.c_double, p64, p64, p64, pi]
    lib.b2_factorize_retry.argtypes = [vp, pd, C.c_double, C.c_double, p64, p64, p64, pi, pi]
    lib.b2_solve.argtypes = [vp, pd, pd, C.c_int, C.c_int, pd]
    lib.b2_solve_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, pd]
    lib.b2_register_host.argtypes = [vp, vp, C.c_size_t]
    lib.b2_unregister_host.argtypes = [vp, vp]
    lib.b2_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.b2_last_timings.argtypes = [vp, pd]
    lib.b2_last_sweeps.argtypes = [vp]
    lib.b2_front_sizes.argtypes = [vp, C.c_int64, p32, p32, p32]
    lib.b2_profile.argtypes = [vp, C.c_int, C.c_int, pi, pi, pi, pd, pi]
    lib.b2_timer_start.argtypes = [vp]
    lib.b2_timer_stop.argtypes = [vp, pd]
    lib.b2_get_perm.argtypes = [vp, p64]
    lib.b2_get_csc.argtypes = [vp, p64, p64]
    lib.b2_get_nzval.argtypes = [vp, pd]
    lib.b2_get_d.argtypes = [vp, pd]
    lib.b2_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    lib.b2_free.argtypes = [vp]
    if hasattr(lib, "b2b_analyze"):
        lib.b2b_analyze.argtypes = [C.c_int64, C.c_int64, p64, p64, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_int64, C.c_int, p64, C.c_int, C.POINTER(vp)]
        lib.b2b_factorize.argtypes = [vp, pd, pu8, C.c_double, p64, p64, p64, p32]
        lib.b2b_refactorize_shift.argtypes = [vp, pd, pd, pu8, C.c_double, p64, p64, p64, p32]
        lib.b2b_solve.argtypes = [vp, pd, pd, pu8, C.c_int]
        lib.b2b_factorize_dev.argtypes = [vp, vp, vp, C.c_double, vp]
        lib.b2b_solve_dev.argtypes = [vp, vp, vp, vp, C.c_int]
        lib.b2b_factor_solve.argtypes = [vp, pd, pd, pd, pu8, C.c_double, C.c_int, p64, p64, p64, p32]
        lib.b2b_factor_solve_dev.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_int, C.c_int, vp]
        lib.b2b_stats.argtypes = [vp, C.POINTER(Stats)]
        lib.b2b_last_ms.argtypes = [vp, pd]
        lib.b2b_timer_start.argtypes = [vp]
        lib.b2b_timer_stop.argtypes = [vp, pd]
        lib.b2b_get_perm.argtypes = [vp, p64]
        lib.b2b_get_d.argtypes = [vp, C.c_int64, pd]
        lib.b2b_free.argtypes = [vp]
    if hasattr(lib, "b2b_nls_dense_solve"):
        lib.b2_nls_default_params.argtypes = [C.POINTER(NLSParams)]
        lib.b2_nls_default_params.restype = None
        lib.b2b_nls_record_len.argtypes = [vp]
        lib.b2b_nls_record_len.restype = C.c_int64
        lib.b2b_nls_dense_solve_dev.argtypes = [vp, C.POINTER(DenseNLS), C.c_int64, C.POINTER(NLSParams), vp, vp]
        lib.b2b_nls_dense_solve.argtypes = [vp, C.POINTER(DenseNLS), C.c_int64, C.POINTER(NLSParams), pd, C.c_int64]
        lib.b2b_nls_dense_submit.argtypes = [vp, C.POINTER(DenseNLS), C.c_int64, C.POINTER(NLSParams), vp, C.c_int]
        lib.b2b_nls_wait.argtypes = [vp]
    if hasattr(lib, "b2_dev_malloc"):
        lib.b2_dev_malloc.argtypes = [C.POINTER(vp), C.c_size_t]
        lib.b2_dev_free.argtypes = [vp]
        lib.b2_dev_upload.argtypes = [vp, vp, C.c_size_t]
        lib.b2_dev_download.argtypes = [vp, vp, C.c_size_t]
    if hasattr(lib, "b2_host_register"):
        lib.b2_host_register.argtypes = [vp, C.c_size_t]
        lib.b2_host_unregister.argtypes = [vp]
    if hasattr(lib, "b2_measure_dgemm"):
        lib.b2_measure_dgemm.argtypes = [C.c_int, C.c_int, pd, pd]
        lib.b2_measure_hbm.argtypes = [C.c_size_t, C.c_int, pd]
    return lib


_LIB = None


def load():
    """The product library.  Raises if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  cannoles_b200 has no CPU fallback.")
        _LIB = bind_library(LIB_PATH)
    return _LIB


def last_error(lib) -> str:
    return lib.b2_last_error().decode("utf-8", "replace")
