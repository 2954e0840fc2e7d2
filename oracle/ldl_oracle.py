"""ctypes face of ``oracle/ldl_oracle.c``: the restated ``LDLFactStruct`` backend.

Mirrors reference/src/solver_types.jl:45-98 (``LDLFactStruct``, ``set_vals!``,
``try_to_factorize``, ``solve_ldl!``, ``get_vals``).  TEST INFRASTRUCTURE ONLY.
PARITY: pinned end-to-end only (reference/test/runtests.jl known answers); the
factorization boundary itself is unpinned in the reference (see the C header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORDER_AMD, ORDER_NATURAL, ORDER_USER = 0, 1, 2


def build_oracle(force: bool = False) -> str:
    """Compile oracle/libldl_oracle.so with gcc (see oracle/Makefile)."""
    so = os.path.join(_HERE, "libldl_oracle.so")
    src = os.path.join(_HERE, "ldl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libldl_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def load_oracle():
    global _LIB
    if _LIB is not None:
        return _LIB
    lib = C.CDLL(build_oracle())
    p64 = C.POINTER(C.c_int64)
    pd = C.POINTER(C.c_double)
    lib.orc_amd.argtypes = [C.c_int64, p64, p64, p64]
    lib.orc_amd.restype = C.c_int
    lib.orc_analyze.argtypes = [C.c_int64, C.c_int64, p64, p64, C.c_int, p64]
    lib.orc_analyze.restype = C.c_void_p
    lib.orc_free.argtypes = [C.c_void_p]
    lib.orc_set_vals.argtypes = [C.c_void_p, pd]
    lib.orc_set_vals_search.argtypes = [C.c_void_p, pd]
    lib.orc_ldl_factorize.argtypes = [C.c_void_p]
    lib.orc_ldl_factorize.restype = C.c_int64
    lib.orc_inertia.argtypes = [C.c_void_p, C.c_double, p64, p64, p64]
    lib.orc_try_to_factorize.argtypes = [C.c_void_p, pd, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_double, C.c_int]
    lib.orc_try_to_factorize.restype = C.c_int
    lib.orc_solve_ldl.argtypes = [C.c_void_p, pd, pd]
    lib.orc_solve_ldl.restype = C.c_int
    lib.orc_matvec.argtypes = [C.c_void_p, pd, pd]
    for name in ("orc_N", "orc_nnzA", "orc_nnzL"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = C.c_int64
    lib.orc_flops.argtypes = [C.c_void_p]
    lib.orc_flops.restype = C.c_double
    for name in ("orc_Ap", "orc_Ai", "orc_slot", "orc_perm", "orc_parent", "orc_Lp"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = p64
    for name in ("orc_Ax", "orc_D"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = pd
    lib.orc_time_factor_solve.argtypes = [C.c_void_p, pd, pd, pd, C.c_int64, C.c_double,
                                          C.c_int, C.c_int, pd, C.POINTER(C.c_int)]
    lib.orc_time_factor_solve.restype = C.c_double
    _LIB = lib
    return lib


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def amd_order(n: int, indptr: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """AMD permutation of a symmetric pattern given as full CSC/CSR (0-based)."""
    lib = load_oracle()
    ap = np.ascontiguousarray(indptr, dtype=np.int64)
    ai = np.ascontiguousarray(indices, dtype=np.int64)
    perm = np.empty(n, dtype=np.int64)
    rc = lib.orc_amd(n, _p64(ap), _p64(ai), _p64(perm))
    if rc != 0:
        raise MemoryError("orc_amd failed")
    return perm


class _Factor:
    """Stand-in for ``LDLFactorizations.LDLFactorization`` (the ``.factor`` field read at
    reference/src/CaNNOLeS.jl:1049)."""

    def __init__(self, owner):
        self._owner = owner

    @property
    def d(self) -> np.ndarray:
        o = self._owner
        return np.ctypeslib.as_array(o._lib.orc_D(o._h), shape=(o.N,)).copy()


class LDLFactStruct:
    """Restated ``LDLFactStruct(N, rows, cols, vals)`` (reference/src/solver_types.jl:61-65).

    ``rows``/``cols`` are the 1-based COO lower triangle, ``vals`` is aliased (not copied), as
    at reference/src/CaNNOLeS.jl:328.
    """

    def __init__(self, N, rows, cols, vals, ordering=ORDER_AMD, perm=None, use_search=False):
        self._lib = load_oracle()
        self.N = int(N)
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.cols = np.ascontiguousarray(cols, dtype=np.int64)
        assert vals.dtype == np.float64 and vals.flags.c_contiguous
        self.vals = vals
        self.use_search = bool(use_search)
        up = None
        if perm is not None:
            ordering = ORDER_USER
            up = np.ascontiguousarray(perm, dtype=np.int64)
        self._h = self._lib.orc_analyze(self.N, len(self.rows), _p64(self.rows), _p64(self.cols),
                                        int(ordering), _p64(up) if up is not None else None)
        if not self._h:
            raise ValueError("orc_analyze: malformed COO (out of range or strictly upper entry)")
        self.factor = _Factor(self)
        self.nfactorize = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.orc_free(h)

    # -- reference verbs -------------------------------------------------------------
    def get_vals(self):
        return self.vals

    def set_vals(self, vals):
        (self._lib.orc_set_vals_search if self.use_search else self._lib.orc_set_vals)(
            self._h, _pd(vals))

    def try_to_factorize(self, vals, nvar, nequ, ncon, eig_tol) -> bool:
        self.nfactorize += 1
        return bool(self._lib.orc_try_to_factorize(self._h, _pd(vals), nvar, nequ, ncon,
                                                   float(eig_tol), int(self.use_search)))

    def solve_ldl(self, rhs, d) -> bool:
        return bool(self._lib.orc_solve_ldl(self._h, _pd(rhs), _pd(d)))

    # -- inspection (tests) ------------------------------------------------------------
    def inertia(self, eig_tol):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._lib.orc_inertia(self._h, float(eig_tol), C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def factorize_only(self) -> int:
        return int(self._lib.orc_ldl_factorize(self._h))

    @property
    def nnzA(self):
        return int(self._lib.orc_nnzA(self._h))

    @property
    def nnzL(self):
        return int(self._lib.orc_nnzL(self._h))

    @property
    def flops(self):
        return float(self._lib.orc_flops(self._h))

    @property
    def colptr(self):
        return np.ctypeslib.as_array(self._lib.orc_Ap(self._h), shape=(self.N + 1,)).copy()

    @property
    def rowval(self):
        return np.ctypeslib.as_array(self._lib.orc_Ai(self._h), shape=(self.nnzA,)).copy()

    @property
    def nzval(self):
        return np.ctypeslib.as_array(self._lib.orc_Ax(self._h), shape=(self.nnzA,)).copy()

    @property
    def slot(self):
        return np.ctypeslib.as_array(self._lib.orc_slot(self._h), shape=(len(self.rows),)).copy()

    @property
    def perm(self):
        return np.ctypeslib.as_array(self._lib.orc_perm(self._h), shape=(self.N,)).copy()

    @property
    def parent(self):
        return np.ctypeslib.as_array(self._lib.orc_parent(self._h), shape=(self.N,)).copy()

    def matvec(self, x):
        y = np.empty(self.N)
        self._lib.orc_matvec(self._h, _pd(np.ascontiguousarray(x, dtype=np.float64)), _pd(y))
        return y

    def time_factor_solve(self, vals, rhs, nvar, eig_tol, reps=1):
        d = np.empty(self.N)
        t3 = np.zeros(3)
        ok = C.c_int(0)
        t = self._lib.orc_time_factor_solve(self._h, _pd(vals), _pd(rhs), _pd(d), nvar,
                                            float(eig_tol), reps, int(self.use_search), _pd(t3),
                                            C.byref(ok))
        return t, t3, bool(ok.value), d
