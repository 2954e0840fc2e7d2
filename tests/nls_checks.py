"""Shared checks for the device-resident batched CaNNOLeS loop (csrc/nls_kernels.cuh): every
instance against the restated per-instance loop (cannoles_b200/solver.py, reference
src/CaNNOLeS.jl:418-864) driven by the CPU oracle on the SAME elimination order.

Tolerances (north_star): iter / nfact / nlinsolve equal; x, objective, ||c|| within 1e-8 relative."""
import functools

import numpy as np

from cannoles_b200.batched_nls import (REC_HEAD, B200BatchNLS, default_params, host_reference_loop,
                                        pack_dense_models, record_dict)
from cannoles_b200.solver import CaNNOLeSSolver, prepare_newton_system
from cannoles_b200.models import DenseBatchNLS

STATUS_CODE = {"first_order": 1, "small_residual": 2, "stalled": 3, "exception": 4, "max_eval": 5,
               "max_time": 6, "max_iter": 7}


def host_loop(oracle_cls, perm, instance, n, m, ncon, x0_scale=1.0, **kw):
    ctor = functools.partial(oracle_cls, perm=perm)
    return host_reference_loop(instance, ctor, n, m, ncon, x0_scale=x0_scale, **kw)


def compare_instance(rec_row, st, nls, n, ncon, tag=""):
    r = record_dict(rec_row, n, ncon)
    ss = st.solver_specific
    assert r["status"] == st.status, (tag, r["status"], st.status)
    assert (r["iter"], r["nfact"], r["nlinsolve"], r["nbk"]) == (st.iter, ss["nfact"], ss["nlinsolve"], ss["nbk"]), \
        (tag, r, st.iter, ss)
    assert (r["neval_residual"], r["neval_cons"]) == (nls.neval_residual, nls.neval_cons), tag
    xs = np.asarray(st.solution)
    assert np.linalg.norm(r["x"] - xs) <= 1e-8 * max(1.0, np.linalg.norm(xs)), (tag, np.linalg.norm(r["x"] - xs))
    assert abs(r["objective"] - st.objective) <= 1e-8 * max(1.0, abs(st.objective)), tag
    assert abs(r["primal_feas"] - st.primal_feas) <= 1e-8 * max(1.0, abs(st.primal_feas)) + 1e-12, tag
    lam = np.asarray(st.multipliers)
    assert np.linalg.norm(r["lam"] - lam) <= 1e-6 * max(1.0, np.linalg.norm(lam)), tag
    return r


def check_first_system_vals(vals_dev, instance, n, m, ncon, x0_scale=1.0):
    """The device fill of the COO values (prepare_newton_system! :947-981) of the FIRST Newton system
    against the host fill at the same (x0, lambda from CGLS, delta)."""
    from cannoles_b200.solver import cgls
    from cannoles_b200.workloads import _NoBackend
    nls = DenseBatchNLS(instance, n, m, ncon)
    s = CaNNOLeSSolver(nls, linsolve=_NoBackend, method="Newton")
    x = nls.x0 * x0_scale
    Fx = np.zeros(m)
    nls.residual(x, Fx)
    nls.jac_coord_residual(x, s.Jx_vals)
    cx = np.zeros(ncon)
    nls.cons(x, cx)
    nls.jac_coord(x, s.Jcx_vals)
    Jxtr = s.Jx.tmul(Fx)
    lam = cgls(s.Jcx, Jxtr)
    if np.linalg.norm(lam) == 0:
        lam[:] = 1.0
    dual = Jxtr - s.Jcx.tmul(lam)
    nd, npr = np.max(np.abs(dual)), np.max(np.abs(cx)) if ncon else 0.0
    delta = max(np.sqrt(2.0 ** -52), min(0.1 * 1.0, nd + npr))
    prepare_newton_system(s, nls, x, lam, Fx, delta)
    ref = s.vals
    # everything but S2 (-0.1 lambda on ncon diagonal entries) is a function of x0 alone; lambda comes
    # out of CGLS stopped at sqrt(eps), so two roundings of it agree to ~1e-8 only
    o2 = s.nnzhF
    seg2 = slice(o2, o2 + s.nnzhc)
    rest = np.ones(len(ref), dtype=bool)
    rest[seg2] = False
    scale = np.max(np.abs(ref[rest]))
    assert np.max(np.abs(vals_dev[rest] - ref[rest])) <= 1e-12 * scale, np.max(np.abs(vals_dev[rest] - ref[rest])) / scale
    if s.nnzhc:
        assert np.max(np.abs(vals_dev[seg2] - ref[seg2])) <= 1e-6 * max(1.0, np.max(np.abs(lam)))


def check_batch_nls(lib, oracle_cls, instances, n=64, m=128, ncon=16, use_host_verb=False, chunk=0, dump_vals=True,
                    params_kw=None, loop_kw=None, x0_scale=1.0):
    instances = list(instances)
    B = len(instances)
    S = B200BatchNLS(B, n, m, ncon, _lib=lib)
    try:
        mod = pack_dense_models(instances, n, m, ncon)
        mod["x0"] *= x0_scale
        prm = default_params(lib, **(params_kw or {}))
        vals = None
        if use_host_verb:
            rec = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"], params=prm, chunk=chunk)
        else:
            ptrs = S.upload(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
            out = S.solve_dev(ptrs, B, params=prm, dump_vals=dump_vals)
            rec, vals = out if dump_vals else (out, None)
        perm = S.kkt.perm
        out = []
        for b, i in enumerate(instances):
            if vals is not None and rec[b, 2] > 0:
                check_first_system_vals(vals[b], i, n, m, ncon, x0_scale)
            st, nls = host_loop(oracle_cls, perm, i, n, m, ncon, x0_scale=x0_scale, **(loop_kw or {}))
            out.append(compare_instance(rec[b], st, nls, n, ncon, tag=f"instance {i}"))
        return rec, out
    finally:
        S.close()
