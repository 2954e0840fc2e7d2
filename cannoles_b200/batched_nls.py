"""Batched ``cannoles`` for config 5: every instance of a batch of small dense constrained NLS
problems is solved to its final status by ONE CTA on the device (csrc/nls_kernels.cuh); the host
sees the model arrays go in and one fixed-size record per instance come out
(reference/src/CaNNOLeS.jl:418-864 per instance; record = the fields of :834-862).

Multi-GPU: ``partition`` the instances over the ranks, each rank calls ``solve`` on
its block, ``batched.gather_records`` collects the records -- the only collective.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import DenseNLS, NLSParams, pd
from .batched import B200BatchStruct
from .linsolve import ORDER_AMD, B200Error
from .models import DenseBatchNLS
from .solver import CaNNOLeSSolver
from .workloads import _NoBackend

STATUS = {0: "unknown", 1: "first_order", 2: "small_residual", 3: "stalled", 4: "exception",
          5: "max_eval", 6: "max_time", 7: "max_iter", 8: "error: Initial point gives Inf or Nan",
          9: "error: Dphi >= 0", 10: "error: alpha too small"}
REC_HEAD = 12
REC_FIELDS = ("status", "iter", "nfact", "nlinsolve", "nbk", "neval_residual", "neval_cons",
              "objective", "primal_feas", "dual_feas", "rho", "delta")


def pack_dense_models(instances, n=64, m=128, ncon=16):
    """Model arrays of ``DenseBatchNLS(i)`` for i in ``instances`` in the layout of
    ``b2_dense_nls_t`` (column-major A, B, C per instance)."""
    instances = list(instances)
    B = len(instances)
    out = {"At": np.empty((B, n, m)), "Bt": np.empty((B, n, m)), "Ct": np.empty((B, n, ncon)),
           "y": np.empty((B, m)), "e": np.empty((B, ncon)), "x0": np.empty((B, n))}
    for b, i in enumerate(instances):
        M = DenseBatchNLS(i, n, m, ncon)
        out["At"][b], out["Bt"][b], out["Ct"][b] = M.A.T, M.B.T, M.C.T
        out["y"][b], out["e"][b], out["x0"][b] = M.y, M.e, M.x0
    return out


def default_params(lib=None, **kw) -> NLSParams:
    lib = lib if lib is not None else _capi.load()
    p = NLSParams()
    lib.b2_nls_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k!r}")
        setattr(p, k, v)
    return p


class B200BatchNLS:
    """``count <= batch`` DenseBatchNLS-shaped instances solved on one GPU."""

    def __init__(self, batch, n=64, m=128, ncon=16, device=0, ordering=ORDER_AMD, perm=None, _lib=None):
        self._lib = _lib if _lib is not None else _capi.load()
        self.n, self.m, self.ncon, self.batch = int(n), int(m), int(ncon), int(batch)
        # the COO layout of the Newton system (src/CaNNOLeS.jl:281-315) of one instance
        s = CaNNOLeSSolver(DenseBatchNLS(0, n, m, ncon), linsolve=_NoBackend, method="Newton")
        self.rows, self.cols = s.rows, s.cols
        self.N = n + m + ncon
        self.kkt = B200BatchStruct(self.N, s.rows, s.cols, batch, n, m, ncon, ordering=ordering, perm=perm,
                                   device=device, _lib=self._lib)
        self.rec_len = int(self._lib.b2b_nls_record_len(self.kkt._h))
        assert self.rec_len == REC_HEAD + n + ncon
        self._dev = []

    def close(self):
        for p in self._dev:
            self._lib.b2_dev_free(p)
        self._dev = []
        self.kkt.close()

    def _check(self, rc):
        if rc != 0:
            raise B200Error(_capi.last_error(self._lib))

    def _model(self, arrs, shared_model):
        md = DenseNLS()
        md.n, md.m, md.ncon, md.shared_model = self.n, self.m, self.ncon, int(bool(shared_model))
        return md

    def solve(self, At, Bt, Ct, y, e, x0, params=None, shared_model=False, chunk=0):
        """Host arrays in, host records out (``count x rec_len``); the model is uploaded in chunks
        that overlap the solves.  Pin the arrays with ``self.kkt.register_host`` for PCIe speed."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (At, Bt, Ct, y, e, x0)]
        count = arrs[5].shape[0]
        md = self._model(arrs, shared_model)
        for name, a in zip(("At", "Bt", "Ct", "y", "e", "x0"), arrs):
            setattr(md, name, a.ctypes.data_as(pd))
        md.y0 = None
        rec = np.zeros((count, self.rec_len))
        prm = params if params is not None else default_params(self._lib)
        self._check(self._lib.b2b_nls_dense_solve(self.kkt._h, C.byref(md), count, C.byref(prm),
                                                  rec.ctypes.data_as(pd), int(chunk)))
        return rec

    def submit(self, arrs, rec, params=None, shared_model=False):
        """Asynchronous ``solve``: queue the batch (host arrays ``arrs`` = (At, Bt, Ct, y, e, x0),
        host records ``rec``) and return; ``wait()`` completes every queued batch.  The arrays must
        stay alive and untouched until then."""
        count = arrs[5].shape[0]
        md = self._model(arrs, shared_model)
        for name, a in zip(("At", "Bt", "Ct", "y", "e", "x0"), arrs):
            assert a.dtype == np.float64 and a.flags.c_contiguous
            setattr(md, name, a.ctypes.data_as(pd))
        md.y0 = None
        prm = params if params is not None else default_params(self._lib)
        self._check(self._lib.b2b_nls_dense_submit(self.kkt._h, C.byref(md), count, C.byref(prm),
                                                   rec.ctypes.data_as(C.c_void_p), 0))

    def submit_dev(self, ptrs, count, drec, params=None, shared_model=False):
        """Asynchronous device-resident form: model pointers from ``upload``, ``drec`` a device buffer."""
        md = self._model(None, shared_model)
        for name, p in zip(("At", "Bt", "Ct", "y", "e", "x0"), ptrs):
            setattr(md, name, C.cast(p, pd))
        md.y0 = None
        prm = params if params is not None else default_params(self._lib)
        self._check(self._lib.b2b_nls_dense_submit(self.kkt._h, C.byref(md), count, C.byref(prm), drec, 1))

    def wait(self):
        self._check(self._lib.b2b_nls_wait(self.kkt._h))

    # -- device-resident model (bench `value` leg, tests) -------------------------------------
    def upload(self, At, Bt, Ct, y, e, x0):
        ptrs = []
        for a in (At, Bt, Ct, y, e, x0):
            a = np.ascontiguousarray(a, dtype=np.float64)
            p = C.c_void_p()
            self._check(self._lib.b2_dev_malloc(C.byref(p), max(a.nbytes, 8)))
            self._check(self._lib.b2_dev_upload(p, a.ctypes.data_as(C.c_void_p), a.nbytes))
            self._dev.append(p)
            ptrs.append(p)
        return ptrs

    def solve_dev(self, ptrs, count, params=None, shared_model=False, dump_vals=False):
        """Model already in HBM (``ptrs`` from ``upload``).  Returns records (and, with
        ``dump_vals``, the COO values of every instance's FIRST Newton system, for parity tests)."""
        md = self._model(None, shared_model)
        for name, p in zip(("At", "Bt", "Ct", "y", "e", "x0"), ptrs):
            setattr(md, name, C.cast(p, pd))
        md.y0 = None
        prm = params if params is not None else default_params(self._lib)
        nrec = count * self.rec_len * 8
        drec, dvals = C.c_void_p(), C.c_void_p()
        self._check(self._lib.b2_dev_malloc(C.byref(drec), nrec))
        nnz = len(self.rows)
        if dump_vals:
            self._check(self._lib.b2_dev_malloc(C.byref(dvals), count * nnz * 8))
        try:
            self._check(self._lib.b2b_nls_dense_solve_dev(self.kkt._h, C.byref(md), count, C.byref(prm), drec,
                                                          dvals if dump_vals else None))
            rec = np.zeros((count, self.rec_len))
            self._check(self._lib.b2_dev_download(rec.ctypes.data_as(C.c_void_p), drec, nrec))
            if dump_vals:
                vals = np.zeros((count, nnz))
                self._check(self._lib.b2_dev_download(vals.ctypes.data_as(C.c_void_p), dvals, vals.nbytes))
                return rec, vals
            return rec
        finally:
            self._lib.b2_dev_free(drec)
            if dump_vals:
                self._lib.b2_dev_free(dvals)

    def last_ms(self):
        return self.kkt.last_ms()


def record_dict(rec_row, n, ncon):
    d = {k: (int(v) if k not in ("objective", "primal_feas", "dual_feas", "rho", "delta") else float(v))
         for k, v in zip(REC_FIELDS, rec_row[:REC_HEAD])}
    d["status"] = STATUS.get(d["status"], "?")
    d["x"] = np.array(rec_row[REC_HEAD:REC_HEAD + n])
    d["lam"] = np.array(rec_row[REC_HEAD + n:REC_HEAD + n + ncon])
    return d


def host_reference_loop(instance, linsolve, n=64, m=128, ncon=16, x0_scale=1.0, **kw):
    """The restated per-instance ``cannoles`` loop (cannoles_b200/solver.py) on one instance with
    the given ``linsolve`` constructor and no wall-clock limit -- what the device loop is compared
    with in the tests."""
    from .solver import solve as _solve
    nls = DenseBatchNLS(instance, n, m, ncon)
    nls.x0 = nls.x0 * x0_scale
    s = CaNNOLeSSolver(nls, linsolve=linsolve, method="Newton")
    st = _solve(s, nls, max_time=math.inf, **kw)
    return st, nls
