"""Host-side caller of the linear-solver boundary: a restatement of CaNNOLeS' ``solve!`` loop.

Julia is not available in this image, so the reference's caller (L3-L5 in SURVEY.md) is
restated here in Python/numpy so that the ``linsolve`` backends can be driven exactly the way
``reference/src/CaNNOLeS.jl`` drives them.  It is NOT accelerated and not part of the hot path:
it exists so that parity (iteration counts, ``nfact``, ``nlinsolve``, final ``x``) can be
measured through the same three verbs the reference calls:

    LDLT = Backend(N, rows, cols, vals)                      src/CaNNOLeS.jl:322-332
    LDLT.try_to_factorize(vals, nvar, nequ, ncon, eig_tol)   src/CaNNOLeS.jl:1023,1032,1039
    LDLT.get_vals()                                          src/CaNNOLeS.jl:1026
    LDLT.solve_ldl(rhs, d)   (d = -K^-1 rhs)                 src/CaNNOLeS.jl:1049

Function-by-function map (all citations into reference/src/CaNNOLeS.jl):
    ParamCaNNOLeS            :36-87
    CaNNOLeSSolver.__init__  :225-377   (COO layout of K: seven segments)
    cannoles                 :402-416
    solve                    :418-864
    optimality_check_small_residual :872-897
    dual_scaling             :917-920
    prepare_newton_system    :947-981  (+ src/hessian_approx.jl:48-60)
    newton_system            :1008-1052
    line_search              :1054-1112
    cgls                     Krylov.jl 0.10 `cgls` [upstream, restated]
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

AVAIL_MTDS = ("Newton", "LM", "Newton_noFHess", "Newton_vanishing")  # :11

# linsolve registry: the product registers "b200"; tests may register the CPU oracle under
# "ldlfactorizations".  Mirrors the if/elseif at :322-332.
LINSOLVE_REGISTRY: dict = {}


def register_linsolve(name: str, ctor) -> None:
    LINSOLVE_REGISTRY[name] = ctor


EPS = float(np.finfo(np.float64).eps)


@dataclass
class ParamCaNNOLeS:
    """:36-87 with T = Float64."""
    eig_tol: float = EPS
    delta_min: float = math.sqrt(EPS)
    kappa_dec: float = 1.0 / 3.0
    kappa_inc: float = 8.0
    kappa_largeinc: float = min(100.0, 8 * 16.0)  # min(100, sizeof(T)*16)
    rho0: float = EPS ** (1.0 / 3.0)
    rho_max: float = min(EPS ** -2.0, float(np.nextafter(np.inf, 0)))
    rho_min: float = math.sqrt(EPS)
    gamma_A: float = EPS ** 0.25


@dataclass
class ExecutionStats:
    """The fields of SolverCore.GenericExecutionStats the reference fills (:604-607, :834-862)."""
    status: str = "unknown"
    solution: np.ndarray | None = None
    objective: float = math.inf
    primal_feas: float = math.inf
    dual_feas: float = math.inf
    multipliers: np.ndarray | None = None
    iter: int = 0
    elapsed_time: float = 0.0
    solver_specific: dict = field(default_factory=dict)
    history: list = field(default_factory=list)  # (iter, inner, rho, delta, nfact) per Newton system


def check_nan_inf(x) -> bool:  # :902-909
    return not bool(np.all(np.isfinite(x)))


def dual_scaling(lam, smax):  # :917-920
    ncon = len(lam)
    return max(smax, np.sum(np.abs(lam)) / ncon) / smax if ncon > 0 else 1.0


class _COO:
    """SparseMatricesCOO stand-in: 1-based rows/cols, shared vals buffer."""

    def __init__(self, m, n, rows, cols, vals):
        self.m, self.n = m, n
        self.r0 = np.asarray(rows, dtype=np.int64) - 1
        self.c0 = np.asarray(cols, dtype=np.int64) - 1
        self.vals = vals

    def tmul(self, v):  # A' v
        if len(self.vals) == 0:
            return np.zeros(self.n)
        return np.bincount(self.c0, weights=self.vals * v[self.r0], minlength=self.n)

    def mul(self, v):  # A v
        if len(self.vals) == 0:
            return np.zeros(self.m)
        return np.bincount(self.r0, weights=self.vals * v[self.c0], minlength=self.m)


def cgls(A: _COO, b, atol=math.sqrt(EPS), rtol=math.sqrt(EPS)):
    """min ||A' x - b|| for x, i.e. Krylov.cgls applied to the operator Jcx' (:513, :887).

    The operator handed to Krylov is ``Jcx'`` (nvar x ncon): products with it are ``A.tmul``,
    products with its adjoint are ``A.mul``.  Stopping rule ||Op' r|| <= atol + rtol ||Op' b||,
    itmax = m + n  [upstream, Krylov 0.10 defaults].
    """
    m, n = A.n, A.m  # operator is n_var x n_con
    x = np.zeros(n)
    bnorm = float(np.linalg.norm(b))
    if bnorm == 0.0 or n == 0:
        return x
    r = b.copy()
    s = A.mul(r)
    p = s.copy()
    gamma = float(s @ s)
    it, itmax = 0, m + n
    arnorm = math.sqrt(gamma)
    eps_ = atol + rtol * arnorm
    solved = arnorm <= eps_
    while not (solved or it >= itmax):
        q = A.tmul(p)
        delta = float(q @ q)
        if delta <= 0.0:
            break
        alpha = gamma / delta
        x += alpha * p
        r -= alpha * q
        s = A.mul(r)
        gamma_next = float(s @ s)
        beta = gamma_next / gamma
        p = s + beta * p
        gamma = gamma_next
        arnorm = math.sqrt(gamma)
        it += 1
        solved = arnorm <= eps_
    return x


def _hess_mode(method, nls):
    """src/hessian_approx.jl:10-40: (nnzhF, rows, cols) of the residual-Hessian segment."""
    if method in ("Newton", "Newton_vanishing"):
        r, c = nls.hess_structure_residual()
        return len(r), np.asarray(r, dtype=np.int64), np.asarray(c, dtype=np.int64)
    return 0, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)


class CaNNOLeSSolver:
    """:181-377.  ``linsolve`` is a registry key ("b200") or a constructor
    ``ctor(N, rows, cols, vals)`` returning an object with the three verbs."""

    def __init__(self, nls, linsolve="b200", method="Newton", linsolve_kwargs=None):
        if method not in AVAIL_MTDS:  # :18-26
            s = "`method` must be one of these: " + ", ".join(f"`{x}`" for x in AVAIL_MTDS)
            raise ValueError(s)
        self.method = method
        nvar, nequ, ncon = nls.nvar, nls.nequ, nls.ncon
        self.nvar, self.nequ, self.ncon = nvar, nequ, ncon
        N = nvar + nequ + ncon
        self.x = np.zeros(nvar)
        self.lam = np.zeros(ncon)
        self.cx = np.zeros(ncon)
        self.r = np.zeros(nequ)
        self.Fx = np.zeros(nequ)
        self.d = np.zeros(N)
        self.dlam = np.zeros(ncon)
        self.rhs = np.zeros(N)
        self.xt = np.zeros(nvar)
        self.rt = np.zeros(nequ)
        self.lamt = np.zeros(ncon)
        self.Ft = np.zeros(nequ)
        self.ct = np.zeros(ncon)

        nnzhc = nls.nnzh if ncon > 0 else 0  # :256
        nnzjF, nnzjc = nls.nnzj_residual, nls.nnzj
        Jx_rows, Jx_cols = nls.jac_structure_residual()
        self.Jx_vals = np.zeros(nnzjF)
        self.Jt_vals = np.zeros(nnzjF)
        self.Jx = _COO(nequ, nvar, Jx_rows, Jx_cols, self.Jx_vals)
        self.Jt = _COO(nequ, nvar, Jx_rows, Jx_cols, self.Jt_vals)
        if ncon > 0:
            Jc_rows, Jc_cols = nls.jac_structure()
        else:
            Jc_rows = Jc_cols = np.zeros(0, dtype=np.int64)
        self.Jcx_vals = np.zeros(nnzjc)
        self.Jct_vals = np.zeros(nnzjc)
        self.Jcx = _COO(ncon, nvar, Jc_rows, Jc_cols, self.Jcx_vals)
        self.Jct = _COO(ncon, nvar, Jc_rows, Jc_cols, self.Jct_vals)

        nnzhF, hr, hc = _hess_mode(method, nls)  # :271-274
        self.nnzhF, self.nnzhc, self.nnzjF, self.nnzjc = nnzhF, nnzhc, nnzjF, nnzjc
        nnzNS = nnzhF + nnzhc + nnzjF + nnzjc + nvar + nequ + ncon  # :273
        rows = np.empty(nnzNS, dtype=np.int64)
        cols = np.empty(nnzNS, dtype=np.int64)
        vals = np.ones(nnzNS)  # :279
        o = 0
        rows[o:o + nnzhF], cols[o:o + nnzhF] = hr, hc  # :284-288
        o += nnzhF
        if ncon > 0:  # :289-292
            r_, c_ = nls.hess_structure()
            rows[o:o + nnzhc], cols[o:o + nnzhc] = r_, c_
        o += nnzhc
        rows[o:o + nnzjF] = np.asarray(Jx_rows, dtype=np.int64) + nvar  # :294-296
        cols[o:o + nnzjF] = Jx_cols
        o += nnzjF
        if ncon > 0:  # :298-302
            rows[o:o + nnzjc] = np.asarray(Jc_rows, dtype=np.int64) + (nvar + nequ)
            cols[o:o + nnzjc] = Jc_cols
        o += nnzjc
        rows[o:o + nequ] = cols[o:o + nequ] = np.arange(nvar + 1, nvar + nequ + 1)  # :304-306
        vals[o:o + nequ] = -1.0
        o += nequ
        if ncon > 0:  # :308-312
            rows[o:o + ncon] = cols[o:o + ncon] = np.arange(nvar + nequ + 1, N + 1)
        o += ncon
        rows[o:o + nvar] = cols[o:o + nvar] = np.arange(1, nvar + 1)  # :314-315
        self.rows, self.cols = rows, cols

        ctor = LINSOLVE_REGISTRY.get(linsolve) if isinstance(linsolve, str) else linsolve
        if ctor is None:  # :330-332
            raise ValueError(f"Can't handle {linsolve}")
        self.LDLT = ctor(N, rows, cols, vals, **(linsolve_kwargs or {}))
        self.vals = self.LDLT.get_vals()  # :328 (alias, no copy)
        self.params = ParamCaNNOLeS()

    def reset(self, nls=None):  # :379-400 (the factor's pattern is NOT re-analysed, as upstream)
        return self


def prepare_newton_system(solver, nls, x, lam, r, delta):
    """:947-981 and src/hessian_approx.jl:48-60."""
    s = solver
    vals = s.vals
    nvar, nequ, ncon = s.nvar, s.nequ, s.ncon
    nnzhF, nnzhc, nnzjF, nnzjc = s.nnzhF, s.nnzhc, s.nnzjF, s.nnzjc
    if s.method in ("Newton", "Newton_vanishing") and nnzhF > 0:
        nls.hess_coord_residual(x, r, vals[0:nnzhF])
    o = nnzhF + nnzhc
    vals[o:o + nnzjF] = s.Jx_vals
    if ncon > 0:
        seg = vals[nnzhF:nnzhF + nnzhc]
        nls.hess_coord(x, lam, seg, obj_weight=0.0)
        np.negative(seg, out=seg)
        o = nnzhF + nnzhc + nnzjF
        vals[o:o + nnzjc] = s.Jcx_vals
        o = nnzhF + nnzhc + nnzjF + nnzjc + nequ
        vals[o:o + ncon] = -delta
    o = nnzhF + nnzhc + nnzjF + nnzjc + nequ + ncon
    vals[o:o + nvar] = 0.0
    return vals


def newton_system(d, nvar, nequ, ncon, rhs, vals, LDLT, rho_old, params):
    """:1008-1052.  Returns (d, solve_success, rho, rho_old, nfact)."""
    nfact = 0
    rho = 0.0
    success = LDLT.try_to_factorize(vals, nvar, nequ, ncon, params.eig_tol)
    nfact += 1
    vals = LDLT.get_vals()
    sI = slice(len(vals) - nvar, len(vals))
    if not success:
        rho = params.rho0 if rho_old == 0 else max(params.rho_min, params.kappa_dec * rho_old)
        vals[sI] = rho
        success = LDLT.try_to_factorize(vals, nvar, nequ, ncon, params.eig_tol)
        nfact += 1
        while (not success) and rho <= params.rho_max:
            rho = params.kappa_largeinc * rho if rho_old == 0 else params.kappa_inc * rho
            if rho <= params.rho_max:
                vals[sI] = rho
                success = LDLT.try_to_factorize(vals, nvar, nequ, ncon, params.eig_tol)
                nfact += 1
        if rho <= params.rho_max:
            rho_old = rho
    solve_success = LDLT.solve_ldl(rhs, d) if success else False
    return d, solve_success, rho, rho_old, nfact


def _phi(lam, Fx, cx, eta):  # :479-481
    return float(Fx @ Fx) / 2 - float(lam @ cx) + eta * float(cx @ cx) / 2


def line_search(solver, nls, x, lam, dx, Fx, cx, Fres, cres, delta, eta, params):
    """:1054-1112 (merit = :auglag, trial_computed = false)."""
    s = solver
    Dphi = float(s.Jx.tmul(Fx) @ dx) - float(dx @ s.Jcx.tmul(lam - cx / delta if len(lam) else lam))
    if len(lam) > 0:
        eta = 1 / delta
    assert Dphi < 0, "Dϕ < 0 violated"  # :1085
    np.add(x, dx, out=s.xt)
    Fres(s.xt, s.Ft)
    cres(s.xt, s.ct)
    phix = _phi(lam, Fx, cx, eta)
    phit = _phi(lam, s.Ft, s.ct, eta)
    alpha = 1.0
    nbk = 0
    while not (phit <= phix + params.gamma_A * alpha * Dphi):
        nbk += 1
        alpha /= 4
        s.xt[:] = x + alpha * dx
        Fres(s.xt, s.Ft)
        cres(s.xt, s.ct)
        phit = _phi(lam, s.Ft, s.ct, eta)
        if alpha < EPS ** 2:
            raise RuntimeError("α too small")
    return eta, alpha, phix, Dphi, nbk


def _get_status(nls, elapsed_time=0.0, iter=0, optimal=False, small_residual=False,
                exception=False, stalled=False, max_eval=math.inf, max_time=math.inf, max_iter=-1):
    """SolverCore.get_status priority order [upstream]."""
    if optimal:
        return "first_order"
    if small_residual:
        return "small_residual"
    if stalled:
        return "stalled"
    if exception:
        return "exception"
    if nls.eval_fun() > max_eval >= 0:
        return "max_eval"
    if elapsed_time > max_time:
        return "max_time"
    if iter > max_iter >= 0:
        return "max_iter"
    return "unknown"


def solve(solver, nls, stats=None, callback=None, x=None, lam=None, use_initial_multiplier=False,
          max_iter=-1, max_eval=100000, max_time=30.0, max_inner=10000, atol=math.sqrt(EPS),
          rtol=math.sqrt(EPS), Fatol=math.sqrt(EPS), Frtol=EPS, verbose=0,
          always_accept_extrapolation=False, delta_dec=0.1):
    """``SolverCore.solve!(solver, nls, stats; ...)`` :418-864."""
    s = solver
    stats = stats if stats is not None else ExecutionStats()
    stats.__init__()
    start_time = time.time()
    nvar, nequ, ncon = s.nvar, s.nequ, s.ncon
    x0 = nls.x0 if x is None else x
    y0 = nls.y0 if lam is None else lam
    s.x[:] = x0
    s.lam[:] = y0
    x, lam = s.x, s.lam
    params = s.params = ParamCaNNOLeS()  # update!(solver.params, eps(T)) :451
    rho = rho_old = 0.0
    delta = 1.0
    vals, LDLT = s.vals, s.LDLT
    Jx, Jcx, Jt, Jct = s.Jx, s.Jcx, s.Jt, s.Jct
    crhs = nls.lcon

    def Fres(x_, Fx_):
        nls.residual(x_, Fx_)
        return Fx_

    def cres(x_, cx_):
        if ncon > 0:
            nls.cons(x_, cx_)
            cx_ -= crhs
        return cx_

    Fx = Fres(x, s.Fx)
    if check_nan_inf(Fx):
        raise RuntimeError("Initial point gives Inf or Nan")
    fx = float(Fx @ Fx) / 2
    nls.jac_coord_residual(x, s.Jx_vals)
    cx = s.cx
    cres(x, cx)
    if ncon > 0:
        nls.jac_coord(x, s.Jcx_vals)
    r = s.r
    r[:] = Fx
    d = s.d
    dx = d[0:nvar]
    dr = d[nvar:nvar + nequ]
    dlam = s.dlam
    Jxtr = Jx.tmul(r)
    if not use_initial_multiplier:
        lam[:] = cgls(Jcx, Jxtr)
        if np.linalg.norm(lam) == 0:
            lam[:] = 1.0
    Jcxtl = Jcx.tmul(lam)
    dual = Jxtr - Jcxtl
    primal = np.concatenate([Fx - r, cx])
    rhs = s.rhs
    normdualhat = normdual = float(np.max(np.abs(dual))) if nvar else 0.0
    normprimalhat = normprimal = float(np.max(np.abs(primal))) if len(primal) else 0.0
    smax = 100.0
    epsF = Fatol + Frtol * 2 * math.sqrt(fx)
    epstol = atol + rtol * normdual
    epsc = math.sqrt(epstol)

    def small_res_check():
        # optimality_check_small_residual! :872-897
        nonlocal Jxtr, Jcxtl, dual
        r[:] = Fx
        Jxtr = Jx.tmul(r)
        lam[:] = cgls(Jcx, Jxtr)
        Jcxtl = Jcx.tmul(lam)
        dual = Jxtr - Jcxtl
        nd = float(np.max(np.abs(dual))) if nvar else 0.0
        primal[0:nequ] = 0.0
        primal[nequ:] = cx
        npz = float(np.max(np.abs(cx))) if ncon else 0.0
        return npz, nd

    small_residual = (2 * math.sqrt(fx) <= epsF) and float(np.linalg.norm(cx)) <= epsc
    sd = dual_scaling(lam, smax)
    first_order = max(normdual / sd, normprimal) <= epstol
    if small_residual and not first_order:
        normprimal, normdual = small_res_check()
        sd = dual_scaling(lam, smax)
        first_order = max(normdual / sd, normprimal) <= epstol
    elapsed = time.time() - start_time
    tired = nls.eval_fun() > max_eval or elapsed > max_time
    broken = False
    internal_msg = ""
    xt, rt, lamt, Ft, ct = s.xt, s.rt, s.lamt, s.Ft, s.ct
    eta = 1.0 if ncon > 0 else 0.0
    stats.iter = 0
    inner_iter = 0
    nbk = nfact = nlinsolve = 0
    epsk = 1e3
    stats.status = _get_status(nls, elapsed_time=elapsed, optimal=first_order,
                               small_residual=small_residual, exception=broken,
                               max_eval=max_eval, max_time=max_time, max_iter=max_iter)
    stats.objective = float(Fx @ Fx) / 2
    stats.primal_feas, stats.dual_feas = float(np.linalg.norm(cx)), normdual
    stats.solution = x.copy()
    stats.multipliers = lam.copy()
    if callback is not None:
        callback(nls, s, stats)
    done = stats.status != "unknown"

    while not done:
        comb = normdual + normprimal
        delta = max(params.delta_min, min(delta_dec * delta, comb))
        inner_iter = 0
        comb_hat = math.inf
        first_iteration = True
        while first_iteration or not (comb_hat <= 0.99 * comb + epsk or tired):
            first_iteration = False
            if inner_iter != 1 or always_accept_extrapolation:
                prepare_newton_system(s, nls, x, lam, r, delta)
                rhs[0:nvar] = dual
                rhs[nvar:] = primal
                d, newton_success, rho, rho_old, nfacti = newton_system(
                    d, nvar, nequ, ncon, rhs, vals, LDLT, rho_old, params)
                nfact += nfacti
                nlinsolve += 1
                stats.history.append((stats.iter, inner_iter, rho, delta, nfacti))
                if rho > params.rho_max or not newton_success or check_nan_inf(d) or fx >= 1e60:
                    if rho > params.rho_max:
                        internal_msg = "ρ → ∞"
                    elif not newton_success:
                        internal_msg = "Failure in Newton step computation"
                    elif np.any(np.isinf(d)):
                        internal_msg = "d → ∞"
                    elif np.any(np.isnan(d)):
                        internal_msg = "d is NaN"
                    else:
                        internal_msg = "f → ∞"
                    broken = True
                    break
                dlam[:] = -d[nvar + nequ:]
            alpha = 0.0
            if inner_iter == 0:
                epsk = max(min(1e3 * delta, 99 * epsk / 100), 9 * epsk / 10)
                np.add(x, dx, out=xt)
                np.add(r, dr, out=rt)
                Mdl = 1e4
                ndl = float(np.linalg.norm(dlam))
                if ndl > Mdl:
                    dlam[:] = dlam * Mdl / ndl
                np.add(lam, dlam, out=lamt)
                Fres(xt, Ft)
                cres(xt, ct)
            else:
                eta, alpha, _phix, _Dphi, nbki = line_search(s, nls, x, lam, dx, Fx, cx, Fres, cres,
                                                            delta, eta, params)
                nbk += nbki
                rt[:] = Ft
                if ncon > 0:
                    lamt[:] = lam - cx / delta
            nls.jac_coord_residual(xt, s.Jt_vals)
            if ncon > 0:
                nls.jac_coord(xt, s.Jct_vals)
            Jxtr = Jt.tmul(rt)
            Jcxtl = Jct.tmul(lamt)
            dual = Jxtr - Jcxtl
            primal[0:nequ] = Ft - rt
            primal[nequ:] = ct
            normdualhat = float(np.max(np.abs(dual))) if nvar else 0.0
            normprimalhat = float(np.max(np.abs(primal))) if len(primal) else 0.0
            comb_hat = normdualhat + normprimalhat
            if inner_iter > 0 or always_accept_extrapolation or comb_hat <= 0.99 * comb + epsk:
                x[:] = xt
                r[:] = rt
                Fx[:] = Ft
                fx = float(Fx @ Fx) / 2
                cx[:] = ct
                s.Jx_vals[:] = s.Jt_vals
                if ncon > 0:
                    s.Jcx_vals[:] = s.Jct_vals
            if comb_hat <= 0.99 * comb + epsk:
                lam[:] = lamt
            else:
                Jxtr = Jx.tmul(r)
                Jcxtl = Jcx.tmul(lam)
                dual = Jxtr - Jcxtl
            if (ncon > 0 and inner_iter > 0 and normdualhat <= 0.99 * normdual + epsk / 2
                    and normprimalhat > 0.99 * normprimal + epsk / 2):
                delta = max(delta / 10, params.delta_min)
            inner_iter += 1
            elapsed = time.time() - start_time
            tired = nls.eval_fun() > max_eval or elapsed > max_time or inner_iter > max_inner
            if verbose > 0 and stats.iter % verbose == 0:
                print(f"{stats.iter:4d} {nls.eval_fun():6d} fx={fx:.3e} dual={normdualhat:.3e} "
                      f"prim={normprimalhat:.3e} a={alpha:.2e} rho={rho:.2e} delta={delta:.2e} "
                      f"in={inner_iter} nbk={nbk}")
        normdual = normdualhat
        normprimal = normprimalhat
        elapsed = time.time() - start_time
        sd = dual_scaling(lam, smax)
        first_order = max(normdual / sd, normprimal) <= epstol
        small_residual = (2 * math.sqrt(fx) <= epsF) and float(np.linalg.norm(cx)) <= epsc
        if small_residual and not first_order:
            normprimal, normdual = small_res_check()
            sd = dual_scaling(lam, smax)
            first_order = max(normdual / sd, normprimal) <= epstol
        stats.iter += 1
        stats.elapsed_time = elapsed
        stats.status = _get_status(nls, elapsed_time=elapsed, iter=stats.iter, optimal=first_order,
                                   small_residual=small_residual, exception=broken,
                                   max_eval=max_eval, max_time=max_time, max_iter=max_iter,
                                   stalled=inner_iter > max_inner >= 0)
        stats.objective = float(Fx @ Fx) / 2
        stats.primal_feas, stats.dual_feas = float(np.linalg.norm(cx)), normdual
        stats.multipliers = lam.copy()
        stats.solution = x.copy()
        if callback is not None:
            callback(nls, s, stats)
        done = stats.status != "unknown"
    stats.solver_specific = {"nbk": nbk, "nfact": nfact, "nlinsolve": nlinsolve,
                             "internal_msg": internal_msg}
    stats.elapsed_time = time.time() - start_time
    return stats


def cannoles(nls, linsolve="b200", method="Newton", linsolve_kwargs=None, **kwargs):
    """:402-416.  (The reference's default ``linsolve`` is ``:ma57``; this build's is ``b200``.)"""
    if nls.has_bounds() or nls.inequality_constrained():
        raise ValueError("Problem has inequalities, can't solve it")
    if not nls.minimize:
        raise ValueError("CaNNOLeS only works for minimization problem")
    solver = CaNNOLeSSolver(nls, linsolve=linsolve, method=method, linsolve_kwargs=linsolve_kwargs)
    return solve(solver, nls, **kwargs)
