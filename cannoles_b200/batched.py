"""Batches of independent KKT systems with one sparsity pattern (BASELINE.json config 5).

``B200BatchStruct`` is the batch counterpart of ``B200Struct``: the same verbs
(reference/src/solver_types.jl:61-98) applied to ``batch`` systems at once, one CTA per
instance (csrc/batched_kernels.cuh), plus ``newton_system`` -- the rho inertia-correction
driver of reference/src/CaNNOLeS.jl:1008-1052 run per instance with masks, so that every
instance sees exactly the rho sequence, ``nfact`` and solution it would see alone.

Multi-GPU: ``partition(batch, rank, world)`` gives the contiguous block of instances of a rank
(SURVEY 8(e)); no data-path collective, one gather of fixed-size per-instance records at the end
(``gather_records``: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _capi
from ._capi import Stats, p32, p64, pd, pu8
from .linsolve import ORDER_AMD, ORDER_USER, B200Error

EPS = 2.0 ** -52


def partition(batch: int, rank: int, world: int):
    """Contiguous block partition: rank g of G owns instances [g*B/G, (g+1)*B/G)."""
    lo = (batch * rank) // world
    hi = (batch * (rank + 1)) // world
    return lo, hi


def _a(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


class B200BatchStruct:
    """``batch`` KKT systems sharing the COO pattern (rows, cols); values are ``batch x nnz``."""

    def __init__(self, N, rows, cols, batch, nvar, nequ, ncon, ordering=ORDER_AMD, perm=None,
                 device=0, _lib=None):
        self._lib = _lib if _lib is not None else _capi.load()
        self.N, self.batch = int(N), int(batch)
        self.nvar, self.nequ, self.ncon = int(nvar), int(nequ), int(ncon)
        self.rows, self.cols = _a(rows, np.int64), _a(cols, np.int64)
        self.nnz = len(self.rows)
        self._h = C.c_void_p()
        up = None
        if perm is not None:
            ordering = ORDER_USER
            self._perm_in = _a(perm, np.int64)
            up = self._perm_in.ctypes.data_as(p64)
        self._check(self._lib.b2b_analyze(self.N, self.nnz, self.rows.ctypes.data_as(p64),
                                          self.cols.ctypes.data_as(p64), self.nvar, self.nequ,
                                          self.ncon, self.batch, int(ordering), up, int(device),
                                          C.byref(self._h)))
        B = self.batch
        self.npos = np.zeros(B, dtype=np.int64)
        self.nzero = np.zeros(B, dtype=np.int64)
        self.nneg = np.zeros(B, dtype=np.int64)
        self.breakdown = np.zeros(B, dtype=np.int32)
        self.nfactorize_calls = 0
        self._pinned = []

    def _check(self, rc):
        if rc != 0:
            raise B200Error(_capi.last_error(self._lib))

    def register_host(self, arr) -> bool:
        """Pin a caller-owned array (values, right-hand sides, steps) so that the H2D / D2H copies of
        the batched verbs run at PCIe speed instead of through a pageable staging buffer."""
        rc = self._lib.b2_host_register(arr.ctypes.data_as(C.c_void_p), arr.nbytes)
        if rc == 0:
            self._pinned.append(arr)
        return rc == 0

    def close(self):
        for a in getattr(self, "_pinned", []):
            self._lib.b2_host_unregister(a.ctypes.data_as(C.c_void_p))
        self._pinned = []
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.b2b_free(h)
            self._h = C.c_void_p()

    __del__ = close

    # ------------------------------------------------------------------------------------
    def _act(self, active):
        if active is None:
            return None, None
        a = _a(active, np.uint8)
        assert a.shape == (self.batch,)
        return a, a.ctypes.data_as(pu8)

    def _ok(self, active):
        ok = (self.npos == self.nvar) & (self.nzero == 0)
        if active is not None:
            ok &= np.asarray(active, dtype=bool)
        return ok

    def _outs(self):
        return (self.npos.ctypes.data_as(p64), self.nzero.ctypes.data_as(p64),
                self.nneg.ctypes.data_as(p64), self.breakdown.ctypes.data_as(p32))

    def try_to_factorize(self, vals, eig_tol=EPS, active=None):
        """Batched ``try_to_factorize``: returns the per-instance success mask."""
        assert vals.dtype == np.float64 and vals.flags.c_contiguous and vals.shape == (self.batch, self.nnz)
        a, pa = self._act(active)
        self._check(self._lib.b2b_factorize(self._h, vals.ctypes.data_as(pd), pa, float(eig_tol), *self._outs()))
        self.nfactorize_calls += 1
        return self._ok(a)

    def refactorize_shift(self, rho, delta=None, eig_tol=EPS, active=None):
        """Retry with per-instance rho (and delta) without re-uploading the values."""
        rho = _a(rho, np.float64)
        dl = None if delta is None else _a(delta, np.float64)
        a, pa = self._act(active)
        self._check(self._lib.b2b_refactorize_shift(self._h, rho.ctypes.data_as(pd),
                                                    None if dl is None else dl.ctypes.data_as(pd), pa,
                                                    float(eig_tol), *self._outs()))
        self.nfactorize_calls += 1
        return self._ok(a)

    def solve_ldl(self, rhs, d, active=None):
        """Batched ``solve_ldl!``: d[b] = -K_b^-1 rhs[b] for the active instances."""
        assert rhs.shape == (self.batch, self.N) and d.shape == (self.batch, self.N)
        assert rhs.flags.c_contiguous and d.flags.c_contiguous
        a, pa = self._act(active)
        self._check(self._lib.b2b_solve(self._h, rhs.ctypes.data_as(pd), d.ctypes.data_as(pd), pa, 1))
        return True

    def factor_solve(self, vals, rhs, d, eig_tol=EPS, active=None):
        """Fused factorize + (where the inertia is right) solve, one kernel launch."""
        assert vals.shape == (self.batch, self.nnz) and rhs.shape == (self.batch, self.N)
        a, pa = self._act(active)
        self._check(self._lib.b2b_factor_solve(self._h, vals.ctypes.data_as(pd), rhs.ctypes.data_as(pd),
                                               d.ctypes.data_as(pd), pa, float(eig_tol), 1, *self._outs()))
        self.nfactorize_calls += 1
        return self._ok(a)

    # -- the rho driver, per instance with masks -----------------------------------------
    def newton_system(self, d, rhs, vals, rho_old, params, active=None):
        """``newton_system!`` (reference/src/CaNNOLeS.jl:1008-1052) for every active instance.

        ``vals`` (batch x nnz) has its rho segment zeroed by the caller (``prepare_newton_system!``
        :978-979).  Returns (solve_success[B], rho[B], rho_old[B], nfact[B]); ``vals``' rho
        segment is left holding the last rho tried, as in the reference."""
        B, nvar = self.batch, self.nvar
        act = np.ones(B, dtype=bool) if active is None else np.asarray(active, dtype=bool).copy()
        rho = np.zeros(B)
        rho_old = np.array(rho_old, dtype=np.float64, copy=True)
        nfact = np.zeros(B, dtype=np.int64)
        success = self.factor_solve(vals, rhs, d, params.eig_tol, act)      # rho = 0 try (+ solve)
        nfact[act] += 1
        solved = success.copy()
        todo = act & ~success
        first = True
        while todo.any():
            if first:
                rho[todo] = np.where(rho_old[todo] == 0, params.rho0,
                                     np.maximum(params.rho_min, params.kappa_dec * rho_old[todo]))
                first = False
                trying = todo.copy()
            else:
                rho[todo] = np.where(rho_old[todo] == 0, params.kappa_largeinc, params.kappa_inc) * rho[todo]
                trying = todo & (rho <= params.rho_max)
            if trying.any():
                vals[trying, self.nnz - nvar:] = rho[trying, None]
                ok = self.refactorize_shift(rho, None, params.eig_tol, trying)
                nfact[trying] += 1
                success |= ok
            todo = todo & ~success & (rho <= params.rho_max)
        upd = act & (rho <= params.rho_max) & (rho != 0)
        # (:1044-1046 sits inside `if !success`; rho stays 0 for first-try successes)
        rho_old[upd] = rho[upd]
        late = success & ~solved
        if late.any():
            self.solve_ldl(rhs, d, late)
        return success, rho, rho_old, nfact

    # -- inspection ------------------------------------------------------------------------
    def stats(self):
        st = Stats()
        self._check(self._lib.b2b_stats(self._h, C.byref(st)))
        return st.as_dict()

    @property
    def perm(self):
        out = np.empty(self.N, dtype=np.int64)
        self._check(self._lib.b2b_get_perm(self._h, out.ctypes.data_as(p64)))
        return out

    def d_of(self, b):
        out = np.empty(self.N)
        self._check(self._lib.b2b_get_d(self._h, int(b), out.ctypes.data_as(pd)))
        return out

    def last_ms(self):
        ms = C.c_double()
        self._check(self._lib.b2b_last_ms(self._h, C.byref(ms)))
        return ms.value


# ------------------------------------------------------------------------------------------
def gather_records(rec: np.ndarray, dist=None, device=None):
    """Gather the fixed-size per-instance records of every rank on all ranks (rank order =
    instance order, because the partition is contiguous).  ``dist`` = torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rec
    import torch
    t = torch.from_numpy(np.ascontiguousarray(rec))
    if device is not None:
        t = t.to(device)
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return np.concatenate([o[:int(s.item())].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)
