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
import os
import time

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

    def _check(self, rc):
        if rc != 0:
            raise B200Error(_capi.last_error(self._lib))

    def close(self):
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


def smoke_batched(batch=6):
    """A few config-5 instances through the fused kernel, checked against the oracle."""
    from oracle import LDLFactStruct
    from .workloads import dense_batch_systems
    s, vals, rhs = dense_batch_systems(range(batch))
    nv, ne, nc = s.nvar, s.nequ, s.ncon
    N = nv + ne + nc
    Bt = B200BatchStruct(N, s.rows, s.cols, batch, nv, ne, nc)
    d = np.zeros((batch, N))
    ok = Bt.factor_solve(vals, rhs, d)
    assert ok.all(), (Bt.npos, Bt.nzero)
    O = LDLFactStruct(N, s.rows, s.cols, vals[0].copy(), perm=Bt.perm)
    for b in range(batch):
        assert O.try_to_factorize(vals[b], nv, ne, nc, EPS)
        assert (Bt.npos[b], Bt.nzero[b], Bt.nneg[b]) == O.inertia(EPS)
        do = np.zeros(N)
        O.solve_ldl(rhs[b], do)
        assert np.linalg.norm(d[b] - do) <= 1e-9 * np.linalg.norm(do)
        r = O.matvec(d[b]) + rhs[b]
        assert np.linalg.norm(r) <= 1e-12 * np.linalg.norm(rhs[b]), np.linalg.norm(r) / np.linalg.norm(rhs[b])
    print("smoke_batched ok: %d instances, N=%d, kernel %.3f ms" % (batch, N, Bt.last_ms()))
    Bt.close()


def bench_batched(args, rank, world, local, dist, total=None, steps=None):
    """C5: `total` independent instances partitioned over the ranks; a step = fused
    factorize + inertia + solve of every instance of the rank (values resident in HBM).
    Returns the "batched" sub-object of the bench line (rank 0) or None."""
    from .workloads import dense_batch_systems
    lib = _capi.load()
    total = int(total or os.environ.get("B2_BENCH_BATCH", 8192))
    steps = int(steps or max(3, min(args.steps, 20)))
    lo, hi = partition(total, rank, world)
    nb = hi - lo
    t0 = time.perf_counter()
    s, vals, rhs = dense_batch_systems(range(lo, hi))
    t_gen = time.perf_counter() - t0
    nv, ne, nc = s.nvar, s.nequ, s.ncon
    N, nnz = nv + ne + nc, vals.shape[1]
    Bt = B200BatchStruct(N, s.rows, s.cols, nb, nv, ne, nc, device=local)
    st = Bt.stats()
    vp = C.c_void_p

    def chk(rc):
        if rc != 0:
            raise B200Error(_capi.last_error(lib))

    dv, dr, do, dc = vp(), vp(), vp(), vp()
    chk(lib.b2_dev_malloc(C.byref(dv), vals.nbytes))
    chk(lib.b2_dev_malloc(C.byref(dr), rhs.nbytes))
    chk(lib.b2_dev_malloc(C.byref(do), rhs.nbytes))
    chk(lib.b2_dev_malloc(C.byref(dc), nb * 32))
    chk(lib.b2_dev_upload(dv, vals.ctypes.data_as(vp), vals.nbytes))
    chk(lib.b2_dev_upload(dr, rhs.ctypes.data_as(vp), rhs.nbytes))
    h = Bt._h

    def sync():
        chk(lib.b2_dev_sync())
        if dist is not None:
            dist.barrier()
        chk(lib.b2_dev_sync())

    for _ in range(3):
        chk(lib.b2b_factor_solve_dev(h, dv, dr, do, None, EPS, 1, 0, dc))
    sync()
    chk(lib.b2b_timer_start(h))
    for _ in range(steps):
        chk(lib.b2b_factor_solve_dev(h, dv, dr, do, None, EPS, 1, 0, dc))
    ms = C.c_double()
    chk(lib.b2b_timer_stop(h, C.byref(ms)))
    sync()
    dev_ms = ms.value
    # end to end from pinned-size host arrays through the public verb
    d = np.zeros((nb, N))
    ok = Bt.factor_solve(vals, rhs, d)
    sync()
    t1 = time.perf_counter()
    chk(lib.b2b_timer_start(h))
    e2e_steps = max(2, steps // 4)
    for _ in range(e2e_steps):
        ok = Bt.factor_solve(vals, rhs, d)
    chk(lib.b2b_timer_stop(h, C.byref(ms)))
    sync()
    e2e_ms = ms.value
    e2e_wall = (time.perf_counter() - t1) * 1e3
    # per-instance record: [ok, npos, nzero, nneg, ||d||]
    rec = np.stack([ok.astype(np.float64), Bt.npos.astype(np.float64), Bt.nzero.astype(np.float64),
                    Bt.nneg.astype(np.float64), np.linalg.norm(d, axis=1)], axis=1)
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0].item()), float(t[1].item())
        allrec = gather_records(rec, dist, device="cuda")
    else:
        allrec = rec
    for p in (dv, dr, do, dc):
        lib.b2_dev_free(p)
    Bt.close()
    if rank != 0:
        return None
    flops_inst = st["flops"]
    out = {"metric": "batched_kkt_factor_solve_per_s", "unit": "instances/s",
           "value": total * steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps,
           "e2e": {"value": total * e2e_steps / (e2e_ms * 1e-3), "unit": "instances/s",
                   "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": e2e_wall / e2e_steps,
                   "h2d_bytes_per_step": int(vals.nbytes + 2 * rhs.nbytes) * world,
                   "d2h_bytes_per_step": int(rhs.nbytes + nb * 32) * world},
           "config": {"workload": "c5: %d independent constrained NLS (n=%d, m=%d, %d constraints), "
                                  "first Newton system of each" % (total, nv, ne, nc),
                      "N": N, "nnz_coo": nnz, "batch_total": total, "batch_per_rank": nb,
                      "nnzL": int(st["nnzL"]), "flops_factor_per_instance": flops_inst,
                      "nsuper": int(st["nsuper"]), "nlevels": int(st["nlevels"])},
           "steps": steps, "n_gpus": world, "scaling": "strong",
           "all_ok": bool((allrec[:, 0] == 1).all()), "records_gathered": int(allrec.shape[0]),
           "roofline": {"bound": "tensor", "unit": "TFLOP/s",
                        "achieved": flops_inst * (total / world) * steps / (dev_ms * 1e-3) / 1e12,
                        "note": "per GPU; algorithmic flops sum_j(c_j^2+3c_j) per instance",
                        "hbm_GBs": (8.0 * nnz + 16.0 * N) * (total / world) * steps / (dev_ms * 1e-3) / 1e9},
           "generate_s": t_gen}
    return out
