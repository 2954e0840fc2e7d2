"""NLS model protocol (the slice of NLPModels.jl that CaNNOLeS calls) and the problem library.

The reference gets its problems from NLPModels/ADNLPModels (reference/src/CaNNOLeS.jl:259-292,
467-477, 490-497, 715-718, 971; src/hessian_approx.jl:25,52).  Julia is absent here, so the
callbacks are restated with numpy.  All structures are 1-based COO, lower triangle for Hessians,
as NLPModels returns them.

Problem library:
  MGH01CON / MGH01_noFHess      reference/test/mgh01con.jl, test/noFHess-model.jl   (config C1)
  SymbolicNLSModel              stands in for ADNLSModel(F, x0, nequ, c, lcon, ucon): sympy
                                derivatives + exact structural sparsity (test/runtests.jl:57-100)
  ExtRosenbrockLinEq            config C2
  BundleAdjustmentLike          config C3
  PoissonParamEst               config C4
  DenseBatchNLS                 config C5 (one instance; ``batch_vals`` builds many)
"""
from __future__ import annotations

import math

import numpy as np


class NLSModel:
    """Equality-constrained NLS model:  min 1/2 ||F(x)||^2  s.t.  c(x) = lcon."""
    minimize = True
    name = "nls"

    def __init__(self, nvar, nequ, ncon, x0, y0=None, lcon=None):
        self.nvar, self.nequ, self.ncon = int(nvar), int(nequ), int(ncon)
        self.x0 = np.array(x0, dtype=np.float64)
        self.y0 = np.zeros(ncon) if y0 is None else np.array(y0, dtype=np.float64)
        self.lcon = np.zeros(ncon) if lcon is None else np.array(lcon, dtype=np.float64)
        self.neval_residual = 0
        self.neval_cons = 0
        # nls_meta.nnzj, nls_meta.nnzh, meta.nnzj, meta.nnzh
        self.nnzj_residual = 0
        self.nnzh_residual = 0
        self.nnzj = 0
        self.nnzh = 0

    # -- NLPModels queries used at src/CaNNOLeS.jl:408-413 -------------------------------
    def has_bounds(self):
        return False

    def inequality_constrained(self):
        return False

    def eval_fun(self):  # SolverCore.eval_fun for NLS models [upstream]
        return self.neval_residual + self.neval_cons

    def reset_counters(self):
        self.neval_residual = self.neval_cons = 0

    # -- to be provided by subclasses ------------------------------------------------------
    def residual(self, x, Fx): raise NotImplementedError
    def jac_structure_residual(self): raise NotImplementedError
    def jac_coord_residual(self, x, vals): raise NotImplementedError
    def hess_structure_residual(self): raise NotImplementedError
    def hess_coord_residual(self, x, v, vals): raise NotImplementedError
    def cons(self, x, cx): raise NotImplementedError
    def jac_structure(self): raise NotImplementedError
    def jac_coord(self, x, vals): raise NotImplementedError
    def hess_structure(self): raise NotImplementedError
    def hess_coord(self, x, y, vals, obj_weight=1.0): raise NotImplementedError


# ------------------------------------------------------------------------------------------
# C1: reference/test/mgh01con.jl
# ------------------------------------------------------------------------------------------
class MGH01CON(NLSModel):
    """Rosenbrock in NLS form with the constraint x1 == 0 (lcon = 0 at test/mgh01con.jl:36)."""
    name = "MGH01CON_manual"

    def __init__(self):
        super().__init__(2, 2, 1, [-1.2, 1.0])
        self.nnzj_residual, self.nnzh_residual, self.nnzj, self.nnzh = 3, 1, 1, 3

    def residual(self, x, Fx):  # :44-50
        self.neval_residual += 1
        Fx[0] = 1 - x[0]
        Fx[1] = 10 * (x[1] - x[0] ** 2)
        return Fx

    def jac_structure_residual(self):  # :53-66
        return np.array([1, 2, 2]), np.array([1, 1, 2])

    def jac_coord_residual(self, x, vals):  # :68-76
        vals[0], vals[1], vals[2] = -1, -20 * x[0], 10
        return vals

    def hess_structure_residual(self):  # :106-115
        return np.array([1]), np.array([1])

    def hess_coord_residual(self, x, v, vals):  # :117-129
        vals[0] = -20 * v[1]
        return vals

    def cons(self, x, cx):  # cons_nln! :193-199
        self.neval_cons += 1
        cx[0] = x[0]
        return cx

    def jac_structure(self):  # :201-210
        return np.array([1]), np.array([1])

    def jac_coord(self, x, vals):  # :212-218
        vals[0] = 1
        return vals

    def hess_structure(self):  # :148-162 (dense lower triangle, column-major)
        return np.array([1, 2, 2]), np.array([1, 1, 2])

    def hess_coord(self, x, y, vals, obj_weight=1.0):  # :258-269 -> :164-177
        vals[0] = 1 - 200 * x[1] + 600 * x[0] ** 2
        vals[1] = -200 * x[0]
        vals[2] = 100
        vals *= obj_weight
        return vals


class MGH01_noFHess(NLSModel):
    """reference/test/noFHess-model.jl: unconstrained Rosenbrock without residual Hessians."""
    name = "MGH01_noFHess"

    def __init__(self):
        super().__init__(2, 2, 0, [-1.2, 1.0])
        self.nnzj_residual, self.nnzh_residual, self.nnzj, self.nnzh = 3, 0, 0, 0

    def residual(self, x, Fx):
        self.neval_residual += 1
        Fx[0] = 1 - x[0]
        Fx[1] = 10 * (x[1] - x[0] ** 2)
        return Fx

    def jac_structure_residual(self):
        return np.array([1, 2, 2]), np.array([1, 1, 2])

    def jac_coord_residual(self, x, vals):
        vals[0], vals[1], vals[2] = -1, -20 * x[0], 10
        return vals

    def hess_structure_residual(self):
        raise TypeError("MethodError: hess_structure_residual not defined for MGH01_noFHess")

    def jac_structure(self):
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)

    def hess_structure(self):
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)


# ------------------------------------------------------------------------------------------
# ADNLSModel stand-in (small problems): symbolic derivatives and exact structural sparsity
# ------------------------------------------------------------------------------------------
class SymbolicNLSModel(NLSModel):
    """``ADNLSModel(F, x0, nequ[, c, lcon, ucon])`` for small problems.

    ``F`` and ``c`` are callables on a list of sympy symbols returning lists of expressions.
    Patterns are the structural nonzeros (what sparse AD detects), column-major like ``findnz``
    of a CSC matrix; the Lagrangian-Hessian pattern is that of
    ``sigma * H(1/2||F||^2) + sum y_i H(c_i)`` (lower triangle), as ``nls.meta.nnzh`` counts it.
    """

    def __init__(self, F, x0, c=None, lcon=None, name="sym"):
        import sympy as sp
        n = len(x0)
        xs = sp.symbols(f"x0:{n}")
        Fe = [sp.sympify(e) for e in F(list(xs))]
        ce = [sp.sympify(e) for e in c(list(xs))] if c is not None else []
        super().__init__(n, len(Fe), len(ce), x0, lcon=lcon)
        self.name = name
        vs = sp.symbols(f"v0:{max(len(Fe), 1)}")
        ys = sp.symbols(f"y0:{max(len(ce), 1)}")
        sig = sp.Symbol("sigma")

        def jac_pattern(exprs):
            ent = []
            for j in range(n):
                for i, e in enumerate(exprs):
                    dij = sp.diff(e, xs[j])
                    if dij != 0:
                        ent.append((i, j, dij))
            return ent

        def lower_pattern(expr):
            ent = []
            for j in range(n):
                gj = sp.diff(expr, xs[j])
                for i in range(j, n):
                    hij = sp.diff(gj, xs[i])
                    if hij != 0:
                        ent.append((i, j, sp.simplify(hij)))
            return ent

        JF = jac_pattern(Fe)
        Jc = jac_pattern(ce)
        HF = lower_pattern(sum(vs[i] * Fe[i] for i in range(len(Fe))))
        obj = sum(e ** 2 for e in Fe) / 2
        lag = sig * obj + sum(ys[i] * ce[i] for i in range(len(ce)))
        HL = lower_pattern(lag)
        self._Fe = sp.lambdify([xs], Fe, "numpy")
        self._ce = sp.lambdify([xs], ce, "numpy") if ce else None
        self._JF = sp.lambdify([xs], [e[2] for e in JF], "numpy")
        self._Jc = sp.lambdify([xs], [e[2] for e in Jc], "numpy")
        self._HF = sp.lambdify([xs, vs], [e[2] for e in HF], "numpy")
        self._HL = sp.lambdify([xs, ys, sig], [e[2] for e in HL], "numpy")
        self._jF = (np.array([e[0] + 1 for e in JF], dtype=np.int64),
                    np.array([e[1] + 1 for e in JF], dtype=np.int64))
        self._jc = (np.array([e[0] + 1 for e in Jc], dtype=np.int64),
                    np.array([e[1] + 1 for e in Jc], dtype=np.int64))
        self._hF = (np.array([e[0] + 1 for e in HF], dtype=np.int64),
                    np.array([e[1] + 1 for e in HF], dtype=np.int64))
        self._hL = (np.array([e[0] + 1 for e in HL], dtype=np.int64),
                    np.array([e[1] + 1 for e in HL], dtype=np.int64))
        self.nnzj_residual, self.nnzh_residual = len(JF), len(HF)
        self.nnzj, self.nnzh = len(Jc), len(HL)
        self._nv, self._ny = len(vs), len(ys)

    def residual(self, x, Fx):
        self.neval_residual += 1
        Fx[:] = np.asarray(self._Fe(x), dtype=np.float64)
        return Fx

    def jac_structure_residual(self):
        return self._jF

    def jac_coord_residual(self, x, vals):
        if len(vals):
            vals[:] = np.asarray(self._JF(x), dtype=np.float64)
        return vals

    def hess_structure_residual(self):
        return self._hF

    def hess_coord_residual(self, x, v, vals):
        if len(vals):
            vv = np.zeros(self._nv)
            vv[:len(v)] = v
            vals[:] = np.asarray(self._HF(x, vv), dtype=np.float64)
        return vals

    def cons(self, x, cx):
        self.neval_cons += 1
        cx[:] = np.asarray(self._ce(x), dtype=np.float64)
        return cx

    def jac_structure(self):
        return self._jc

    def jac_coord(self, x, vals):
        if len(vals):
            vals[:] = np.asarray(self._Jc(x), dtype=np.float64)
        return vals

    def hess_structure(self):
        return self._hL

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        if len(vals):
            yy = np.zeros(self._ny)
            yy[:len(y)] = y
            vals[:] = np.asarray(self._HL(x, yy, obj_weight), dtype=np.float64)
        return vals


# ------------------------------------------------------------------------------------------
# C2: extended Rosenbrock + linear equality block (Gauss-Newton Hessian)
# ------------------------------------------------------------------------------------------
class ExtRosenbrockLinEq(NLSModel):
    """F = F_larger of reference/test/runtests.jl:59 at size n; c_j = x_{2j-1} + x_{2j} - 2.

    nvar n, nequ 2(n-1), ncon n/2.  Lagrangian-Hessian pattern: the AD-like tridiagonal
    (n + n-1 entries) whose constraint part is numerically zero (linear constraints); with
    ``lean=True`` the pattern is empty.  Intended for ``method="Newton_noFHess"``.
    """
    name = "ext_rosenbrock_lineq"

    def __init__(self, n=100_000, seed=1, lean=False):
        assert n % 2 == 0
        rng = np.random.default_rng(seed)
        x0 = 0.9 * np.ones(n) + 0.01 * rng.standard_normal(n)
        super().__init__(n, 2 * (n - 1), n // 2, x0, lcon=np.zeros(n // 2))
        self.n = n
        i = np.arange(1, n)  # 1..n-1
        # rows i: d/dx_i = -20 x_i, d/dx_{i+1} = 10 ; rows n-1+i: d/dx_i = 1
        self._jr = np.concatenate([i, i, (n - 1) + i]).astype(np.int64)
        self._jc_ = np.concatenate([i, i + 1, i]).astype(np.int64)
        j = np.arange(1, n // 2 + 1)
        self._cr = np.concatenate([j, j]).astype(np.int64)
        self._cc = np.concatenate([2 * j - 1, 2 * j]).astype(np.int64)
        if lean:
            self._hr = self._hc = np.zeros(0, dtype=np.int64)
        else:
            k = np.arange(1, n + 1)
            self._hr = np.concatenate([k, k[1:]]).astype(np.int64)
            self._hc = np.concatenate([k, k[:-1]]).astype(np.int64)
        self.nnzj_residual, self.nnzh_residual = 3 * (n - 1), n - 1
        self.nnzj, self.nnzh = n, len(self._hr)

    def residual(self, x, Fx):
        self.neval_residual += 1
        n = self.n
        Fx[0:n - 1] = 10 * (x[1:] - x[:-1] ** 2)
        Fx[n - 1:] = x[:-1] - 1
        return Fx

    def jac_structure_residual(self):
        return self._jr, self._jc_

    def jac_coord_residual(self, x, vals):
        n = self.n
        vals[0:n - 1] = -20 * x[:-1]
        vals[n - 1:2 * (n - 1)] = 10
        vals[2 * (n - 1):] = 1
        return vals

    def hess_structure_residual(self):
        k = np.arange(1, self.n)
        return k, k

    def hess_coord_residual(self, x, v, vals):
        vals[:] = -20 * v[0:self.n - 1]
        return vals

    def cons(self, x, cx):
        self.neval_cons += 1
        cx[:] = x[0::2] + x[1::2] - 2
        return cx

    def jac_structure(self):
        return self._cr, self._cc

    def jac_coord(self, x, vals):
        vals[:] = 1
        return vals

    def hess_structure(self):
        return self._hr, self._hc

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        vals[:] = 0.0
        if obj_weight != 0.0 and len(vals):
            raise NotImplementedError("objective Hessian not needed by CaNNOLeS (obj_weight=0)")
        return vals


# ------------------------------------------------------------------------------------------
# C4: 2-D Poisson-constrained parameter estimation
# ------------------------------------------------------------------------------------------
class PoissonParamEst(NLSModel):
    """Unknowns (u, q) on a g x g grid (G = g^2 each).

    F = [u_i^2 - d_i ; sqrt(alpha) (q_i - qbar_i)]          (nequ = 2G)
    c_i = (A u)_i + q_i u_i - f_i,  A = 5-point Laplacian, Dirichlet   (ncon = G)
    Exact residual Hessian (``method="Newton"``): diag(2 r_i) on u (nnzhF = G);
    Lagrangian-Hessian pattern: the cross terms (q_i, u_i) with value y_i (nnzhc = G).
    """
    name = "poisson_param_est"

    def __init__(self, g=512, alpha=1e-3, seed=3):
        G = g * g
        rng = np.random.default_rng(seed)
        u0 = 1 + 0.1 * rng.standard_normal(G)
        q0 = np.ones(G)
        super().__init__(2 * G, 2 * G, G, np.concatenate([u0, q0]), lcon=np.zeros(G))
        self.g, self.G, self.sa = g, G, math.sqrt(alpha)
        ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
        xs, ys = (ii.ravel() + 1) / (g + 1), (jj.ravel() + 1) / (g + 1)
        utrue = 1 + 0.5 * np.sin(math.pi * xs) * np.sin(math.pi * ys)
        self.dobs = utrue ** 2 * (1 + 0.01 * rng.standard_normal(G))
        self.qbar = np.ones(G)
        idx = np.arange(G)
        rows, cols, vals = [idx], [idx], [4.0 * np.ones(G)]
        for (di, dj) in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            ok = ((ii + di >= 0) & (ii + di < g) & (jj + dj >= 0) & (jj + dj < g)).ravel()
            rows.append(idx[ok])
            cols.append(((ii + di) * g + (jj + dj)).ravel()[ok])
            vals.append(-np.ones(ok.sum()))
        self._Ar = np.concatenate(rows)
        self._Ac = np.concatenate(cols)
        self._Av = np.concatenate(vals) * (g + 1) ** 0  # unscaled stencil keeps entries O(1)
        self._ndiag = G
        qtrue = 1 + 0.2 * np.cos(math.pi * xs)
        self.f = self._Amul(utrue) + qtrue * utrue
        self.nnzj_residual, self.nnzh_residual = 2 * G, G
        self.nnzj, self.nnzh = len(self._Ar) + G, G

    def _Amul(self, u):
        return np.bincount(self._Ar, weights=self._Av * u[self._Ac], minlength=self.G)

    def residual(self, x, Fx):
        self.neval_residual += 1
        G = self.G
        Fx[:G] = x[:G] ** 2 - self.dobs
        Fx[G:] = self.sa * (x[G:] - self.qbar)
        return Fx

    def jac_structure_residual(self):
        k = np.arange(1, 2 * self.G + 1)
        return k, k

    def jac_coord_residual(self, x, vals):
        G = self.G
        vals[:G] = 2 * x[:G]
        vals[G:] = self.sa
        return vals

    def hess_structure_residual(self):
        k = np.arange(1, self.G + 1)
        return k, k

    def hess_coord_residual(self, x, v, vals):
        vals[:] = 2 * v[:self.G]
        return vals

    def cons(self, x, cx):
        self.neval_cons += 1
        G = self.G
        cx[:] = self._Amul(x[:G]) + x[G:] * x[:G] - self.f
        return cx

    def jac_structure(self):
        G = self.G
        k = np.arange(1, G + 1)
        return (np.concatenate([self._Ar + 1, k]).astype(np.int64),
                np.concatenate([self._Ac + 1, G + k]).astype(np.int64))

    def jac_coord(self, x, vals):
        G = self.G
        vals[:len(self._Av)] = self._Av
        vals[:G] += x[G:]            # the first G stencil entries are the diagonal
        vals[len(self._Av):] = x[:G]
        return vals

    def hess_structure(self):
        k = np.arange(1, self.G + 1)
        return self.G + k, k

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        if obj_weight != 0.0:
            raise NotImplementedError("objective Hessian not needed by CaNNOLeS (obj_weight=0)")
        vals[:] = y
        return vals


# ------------------------------------------------------------------------------------------
# C3: bundle-adjustment-shaped NLS (Gauss-Newton)
# ------------------------------------------------------------------------------------------
class BundleAdjustmentLike(NLSModel):
    """ncam cameras x 9 parameters, npts points x 3; every point is seen by two cameras
    (``floor(p/ppc)`` and one of the next 8), two residual rows per observation, 7 gauge
    constraints fixing the first 7 camera parameters.  Residual rows are affine-plus-bilinear
    in (camera, point) so that the Jacobian has the 9+3 block shape of a reprojection error:

        F_o = W_o [cam_o ; pt_o] + 0.1 * (cam_o[0:2] * pt_o[0:2]) - b_o      (2 rows)

    Intended for ``method="Newton_noFHess"``.
    """
    name = "bundle_adjustment_like"

    def __init__(self, ncam=50_000, npts=1_000_000, seed=2):
        rng = np.random.default_rng(seed)
        nvar = 9 * ncam + 3 * npts
        nobs = 2 * npts
        ppc = max(1, npts // ncam)
        p = np.arange(npts)
        c1 = np.minimum(p // ppc, ncam - 1)
        c2 = (c1 + 1 + rng.integers(0, 8, size=npts)) % ncam
        self.obs_cam = np.stack([c1, c2], axis=1).ravel()       # (nobs,)
        self.obs_pt = np.repeat(p, 2)
        self.ncam, self.npts, self.nobs = ncam, npts, nobs
        self.W = rng.standard_normal((nobs, 2, 12))
        self.W[:, :, :9] *= 1.0
        self.W[:, :, 9:] *= 0.3
        x0 = np.concatenate([rng.standard_normal(9 * ncam), rng.standard_normal(3 * npts)])
        super().__init__(nvar, 2 * nobs, 7, x0, lcon=x0[:7].copy())
        xt = x0 + 0.05 * rng.standard_normal(nvar)
        self.b = np.zeros(2 * nobs)
        self.b = self._F(xt)
        o = np.arange(nobs)
        cidx = 9 * self.obs_cam[:, None] + np.arange(9)[None, :]            # (nobs, 9)
        pidx = 9 * ncam + 3 * self.obs_pt[:, None] + np.arange(3)[None, :]  # (nobs, 3)
        self._vidx = np.concatenate([cidx, pidx], axis=1)                   # (nobs, 12)
        rows = (2 * o[:, None, None] + np.arange(2)[None, :, None]) + np.zeros((1, 1, 12), int)
        cols = self._vidx[:, None, :] + np.zeros((1, 2, 1), int)
        self._jr = (rows.ravel() + 1).astype(np.int64)
        self._jc_ = (cols.ravel() + 1).astype(np.int64)
        self.nnzj_residual, self.nnzh_residual = 24 * nobs, 0
        self.nnzj, self.nnzh = 7, 0

    def _F(self, x):
        z = x[self._vidx] if hasattr(self, "_vidx") else None
        if z is None:
            cidx = 9 * self.obs_cam[:, None] + np.arange(9)[None, :]
            pidx = 9 * self.ncam + 3 * self.obs_pt[:, None] + np.arange(3)[None, :]
            z = x[np.concatenate([cidx, pidx], axis=1)]
        lin = np.einsum("okj,oj->ok", self.W, z)
        lin += 0.1 * z[:, 0:2] * z[:, 9:11]
        return lin.ravel() - self.b

    def residual(self, x, Fx):
        self.neval_residual += 1
        Fx[:] = self._F(x)
        return Fx

    def jac_structure_residual(self):
        return self._jr, self._jc_

    def jac_coord_residual(self, x, vals):
        z = x[self._vidx]
        J = self.W.copy()
        for k in range(2):
            J[:, k, k] += 0.1 * z[:, 9 + k]
            J[:, k, 9 + k] += 0.1 * z[:, k]
        vals[:] = J.ravel()
        return vals

    def hess_structure_residual(self):
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)

    def cons(self, x, cx):
        self.neval_cons += 1
        cx[:] = x[:7]
        return cx

    def jac_structure(self):
        k = np.arange(1, 8)
        return k, k

    def jac_coord(self, x, vals):
        vals[:] = 1
        return vals

    def hess_structure(self):
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        return vals


# ------------------------------------------------------------------------------------------
# C5: small dense constrained NLS (one instance of the batch)
# ------------------------------------------------------------------------------------------
class DenseBatchNLS(NLSModel):
    """F(x) = A x + 0.1 sin(B x) - y ;  c(x) = C x + 0.05 (x.x)[0:ncon] - e.

    n = 64, m = 128, 16 constraints; A, B, C ~ N(0,1)/sqrt(n); per-instance seed 1000 + i.
    Dense J (m n), dense Jc (ncon n), dense lower-triangular Hessian patterns (n(n+1)/2 each).
    """
    name = "dense_batch_nls"

    def __init__(self, instance=0, n=64, m=128, ncon=16):
        rng = np.random.default_rng(1000 + instance)
        self.A = rng.standard_normal((m, n)) / math.sqrt(n)
        self.B = rng.standard_normal((m, n)) / math.sqrt(n)
        self.C = rng.standard_normal((ncon, n)) / math.sqrt(n)
        xs = rng.standard_normal(n)
        self.y = self.A @ xs + 0.1 * np.sin(self.B @ xs) + 0.01 * rng.standard_normal(m)
        self.e = self.C @ xs + 0.05 * (xs * xs)[:ncon]
        x0 = rng.standard_normal(n)
        super().__init__(n, m, ncon, x0, lcon=np.zeros(ncon))
        self.n, self.m = n, m
        jj, ii = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")   # column-major
        self._jr, self._jc_ = (ii.ravel() + 1).astype(np.int64), (jj.ravel() + 1).astype(np.int64)
        jj, ii = np.meshgrid(np.arange(n), np.arange(ncon), indexing="ij")
        self._cr, self._cc = (ii.ravel() + 1).astype(np.int64), (jj.ravel() + 1).astype(np.int64)
        lr, lc = [], []
        for j in range(n):
            lr.append(np.arange(j, n))
            lc.append(np.full(n - j, j))
        self._lr = (np.concatenate(lr) + 1).astype(np.int64)
        self._lc = (np.concatenate(lc) + 1).astype(np.int64)
        self.nnzj_residual = m * n
        self.nnzh_residual = n * (n + 1) // 2
        self.nnzj = ncon * n
        self.nnzh = n * (n + 1) // 2

    def residual(self, x, Fx):
        self.neval_residual += 1
        Fx[:] = self.A @ x + 0.1 * np.sin(self.B @ x) - self.y
        return Fx

    def jac_structure_residual(self):
        return self._jr, self._jc_

    def _J(self, x):
        return self.A + 0.1 * np.cos(self.B @ x)[:, None] * self.B

    def jac_coord_residual(self, x, vals):
        vals[:] = self._J(x).T.ravel()   # column-major
        return vals

    def hess_structure_residual(self):
        return self._lr, self._lc

    def hess_coord_residual(self, x, v, vals):
        w = -0.1 * np.sin(self.B @ x) * v
        H = (self.B * w[:, None]).T @ self.B
        vals[:] = H[self._lr - 1, self._lc - 1]
        return vals

    def cons(self, x, cx):
        self.neval_cons += 1
        cx[:] = self.C @ x + 0.05 * (x * x)[:self.ncon] - self.e
        return cx

    def jac_structure(self):
        return self._cr, self._cc

    def jac_coord(self, x, vals):
        Jc = self.C.copy()
        k = np.arange(self.ncon)
        Jc[k, k] += 0.1 * x[:self.ncon]
        vals[:] = Jc.T.ravel()
        return vals

    def hess_structure(self):
        return self._lr, self._lc

    def hess_coord(self, x, y, vals, obj_weight=1.0):
        if obj_weight != 0.0:
            raise NotImplementedError("objective Hessian not needed by CaNNOLeS (obj_weight=0)")
        vals[:] = 0.0
        diag = self._lr == self._lc
        dv = np.zeros(self.n)
        dv[:self.ncon] = 0.1 * y
        vals[diag] = dv
        return vals
