"""`linsolve = :b200` -- the host-side mirror of the reference's linear-solver struct interface.

Mirrors reference/src/solver_types.jl with the same names and argument meaning:

    B200Struct(N, rows, cols, vals)                               ctor        :61-65
    get_vals(LDLT)                                                :67
    try_to_factorize(LDLT, vals, nvar, nequ, ncon, eig_tol)::Bool :79-98
    solve_ldl!(rhs, LDLT.factor, d)::Bool   (d = -K^-1 rhs)        :69-77

Everything numerical happens in libcannoles_b200.so (CUDA, sm_100a) through the C ABI of
include/cannoles_b200.h; this file only marshals pointers.  Numerical failure returns False and
never raises (as :41-42, :96-97); CUDA/runtime errors raise ``B200Error``.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import Stats, p64, pd

ORDER_ND, ORDER_NATURAL, ORDER_USER, ORDER_AMD = 0, 1, 2, 3


class B200Error(RuntimeError):
    pass


def _p64(a):
    return a.ctypes.data_as(p64)


def _pd(a):
    return a.ctypes.data_as(pd)


class B200Factor:
    """What ``LDLT.factor`` is for this backend (read at reference/src/CaNNOLeS.jl:1049)."""

    def __init__(self, owner):
        self._owner = owner

    @property
    def d(self) -> np.ndarray:
        """Pivots in elimination order (the ``factor.d`` the reference's inertia loop reads)."""
        o = self._owner
        out = np.empty(o.N)
        o._check(o._lib.b2_get_d(o._h, _pd(out)))
        return out


class B200Struct:
    """``B200Struct <: LinearSolverStruct``.

    ``rows``/``cols``: 1-based COO lower triangle of K (duplicates allowed); ``vals`` is ALIASED,
    not copied, exactly like ``LDLFactStruct`` (reference/src/CaNNOLeS.jl:328).  ``nvar``/
    ``nequ``/``ncon`` are optional at construction (the reference's ctor does not receive them);
    when absent they are inferred at the first ``try_to_factorize`` and the symbolic analysis is
    done then.

    ``refine_steps`` is the MAXIMUM number of iterative-refinement sweeps of ``solve_ldl``; a sweep
    is taken only while ||K d + rhs|| / ||rhs|| > ``refine_tol`` (north_star bar: 1e-12; the default 5e-13
    keeps a factor 2 below it -- the residual is the true one, formed on the device after every sweep;
    1e-13 cost config 3 a second 9 ms sweep for a first-sweep residual of 1.1e-13).

    ``shift_retries=True`` exploits the caller protocol of ``newton_system!``
    (reference/src/CaNNOLeS.jl:1023-1043): a call whose trailing rho segment is a non-zero
    constant is *probably* a retry of the previous call with only that segment changed, so the
    diagonal is shifted on the device and the factorization starts at once (bit-identical CSC
    values) while the upload and a device-side comparison with the previous values run on a copy
    stream; if anything else changed, the full factorization of the new values is what is returned
    (``b2_factorize_retry``).  Correctness never depends on the caller following the protocol.
    """

    def __init__(self, N, rows, cols, vals, nvar=None, nequ=None, ncon=None, ordering=ORDER_ND,
                 perm=None, device=0, refine_steps=1, refine_tol=5e-13, shift_retries=True, pin=True,
                 _lib=None):
        self._lib = _lib if _lib is not None else _capi.load()
        self.N = int(N)
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.cols = np.ascontiguousarray(cols, dtype=np.int64)
        if not (isinstance(vals, np.ndarray) and vals.dtype == np.float64 and vals.flags.c_contiguous):
            raise TypeError("vals must be a contiguous float64 array (it is aliased, not copied)")
        if len(self.rows) != len(vals) or len(self.cols) != len(vals):
            raise ValueError("rows, cols, vals must have equal length")
        self.vals = vals
        self.ordering = ORDER_USER if perm is not None else int(ordering)
        self._perm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int64)
        self.device = int(device)
        self.refine_steps = int(refine_steps)
        self.refine_tol = float(refine_tol)
        self.shift_retries = bool(shift_retries)
        self.pin = bool(pin)
        self._h = C.c_void_p()
        self._pinned = []
        self.factor = B200Factor(self)
        self.last_inertia = None
        self.last_relres = None
        self.n_upload = 0
        self.n_shift = 0
        self._dims = None
        if nvar is not None:
            self._analyze(int(nvar), int(nequ), int(ncon))

    # ------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise B200Error(_capi.last_error(self._lib))

    def _analyze(self, nvar, nequ, ncon):
        if nvar + nequ + ncon != self.N:
            raise ValueError("nvar + nequ + ncon != N")
        up = _p64(self._perm) if self._perm is not None else None
        self._check(self._lib.b2_analyze(self.N, len(self.vals), _p64(self.rows), _p64(self.cols),
                                         nvar, nequ, ncon, self.ordering, up, self.device,
                                         C.byref(self._h)))
        self._dims = (nvar, nequ, ncon)
        self._check(self._lib.b2_set_option(self._h, b"refine_tol", self.refine_tol))
        if self.pin:
            self.register_host(self.vals)

    def register_host(self, arr):
        """Pin a caller-owned buffer (vals / rhs / d have stable addresses in the reference)."""
        rc = self._lib.b2_register_host(self._h, arr.ctypes.data_as(C.c_void_p), arr.nbytes)
        if rc == 0:
            self._pinned.append(arr)
        return rc == 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.b2_free(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    close = __del__

    # -- the reference's verbs -------------------------------------------------------------
    def get_vals(self):
        return self.vals

    def try_to_factorize(self, vals, nvar, nequ, ncon, eig_tol) -> bool:
        if self._dims is None:
            self._analyze(int(nvar), int(nequ), int(ncon))
        elif self._dims != (nvar, nequ, ncon):
            raise ValueError("block sizes differ from the analysed ones")
        if vals is not self.vals and not np.shares_memory(vals, self.vals):
            if vals.dtype != np.float64 or len(vals) != len(self.vals):
                raise TypeError("vals must be a float64 array of length nnz")
            vals = np.ascontiguousarray(vals)
        npos, nzero, nneg = C.c_int64(), C.c_int64(), C.c_int64()
        brk = C.c_int()
        n = len(vals)
        rho = vals[n - 1] if nvar > 0 else 0.0
        # A call that LOOKS like a rho retry of newton_system! (trailing segment a non-zero constant,
        # reference/src/CaNNOLeS.jl:1029-1043) goes through b2_factorize_retry: the shifted matrix is
        # factorized speculatively while the device checks that nothing else changed since the last
        # upload -- a caller that also edited H/J gets the full factorization of what it passed.
        looks_like_retry = (self.shift_retries and self.n_upload > 0 and nvar > 0 and rho != 0.0
                            and vals[n - nvar] == rho)
        if looks_like_retry:
            held = C.c_int()
            self._check(self._lib.b2_factorize_retry(self._h, _pd(vals), float(rho), float(eig_tol),
                                                     C.byref(npos), C.byref(nzero), C.byref(nneg),
                                                     C.byref(brk), C.byref(held)))
            if held.value:
                self.n_shift += 1
            else:
                self.n_upload += 1
        else:
            self._check(self._lib.b2_factorize(self._h, _pd(vals), float(eig_tol), C.byref(npos),
                                               C.byref(nzero), C.byref(nneg), C.byref(brk)))
            self.n_upload += 1
        self.last_inertia = (npos.value, nzero.value, nneg.value, bool(brk.value))
        return npos.value == nvar and nzero.value == 0

    def solve_ldl(self, rhs, d) -> bool:
        """``solve_ldl!(rhs, LDLT.factor, d)``: d = -(K^-1 rhs); returns True (as the reference)."""
        rr = C.c_double()
        self._check(self._lib.b2_solve(self._h, _pd(rhs), _pd(d), 1, self.refine_steps, C.byref(rr)))
        self.last_relres = rr.value
        return True

    # -- inspection ------------------------------------------------------------------------
    def stats(self) -> dict:
        st = Stats()
        self._check(self._lib.b2_stats(self._h, C.byref(st)))
        return st.as_dict()

    def timings(self) -> dict:
        ms = np.zeros(5)
        self._check(self._lib.b2_last_timings(self._h, _pd(ms)))
        return dict(zip(("upload", "assemble", "factor", "solve", "download"), ms.tolist()))

    @property
    def last_sweeps(self) -> int:
        """Forward/backward sweeps of the last solve (1 = no refinement was needed)."""
        return int(self._lib.b2_last_sweeps(self._h))

    @property
    def perm(self):
        out = np.empty(self.N, dtype=np.int64)
        self._check(self._lib.b2_get_perm(self._h, _p64(out)))
        return out

    @property
    def nzval(self):
        st = self.stats()
        out = np.empty(st["nnzA"])
        self._check(self._lib.b2_get_nzval(self._h, _pd(out)))
        return out

    def csc(self):
        st = self.stats()
        cp = np.empty(self.N + 1, dtype=np.int64)
        rv = np.empty(st["nnzA"], dtype=np.int64)
        self._check(self._lib.b2_get_csc(self._h, _p64(cp), _p64(rv)))
        return cp, rv
