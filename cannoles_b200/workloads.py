"""Synthetic instances of the BASELINE.json configs at a fixed iterate.

``first_system`` builds the restated ``CaNNOLeSSolver`` for a model with a given ``linsolve``
constructor and fills the COO values / right-hand side of its FIRST Newton system exactly as
``prepare_newton_system!`` does (reference/src/CaNNOLeS.jl:947-981), so that the three backend
verbs can be driven (and timed) without running the optimisation loop.
"""
from __future__ import annotations

import numpy as np

from .models import (BundleAdjustmentLike, DenseBatchNLS, ExtRosenbrockLinEq, MGH01CON,
                     PoissonParamEst)
from .solver import CaNNOLeSSolver, prepare_newton_system

EPS = 2.0 ** -52


def make_config(name: str, size: int | None = None):
    """(model, hessian mode, description) for the named BASELINE.json config."""
    if name == "c1":
        return MGH01CON(), "Newton", "mgh01con (n=2, 1 constraint)"
    if name == "c2":
        n = size or 100_000
        return ExtRosenbrockLinEq(n), "Newton_noFHess", f"extended Rosenbrock n={n} + linear equalities (GN)"
    if name == "c3":
        ncam = size or 50_000
        return (BundleAdjustmentLike(ncam, 20 * ncam), "Newton_noFHess",
                f"bundle-adjustment-shaped NLS: {ncam} cameras, {20 * ncam} points (GN)")
    if name == "c4":
        g = size or 512
        return PoissonParamEst(g), "Newton", f"2D Poisson-constrained parameter estimation {g}x{g} (exact residual Hessian)"
    if name == "c5":
        return DenseBatchNLS(0), "Newton", "dense NLS n=64 m=128 16 constraints (one instance of the batch)"
    raise ValueError(f"unknown config {name!r}")


class _NoBackend:
    """Placeholder ``linsolve`` for building only the COO structure of a solver."""

    def __init__(self, N, rows, cols, vals):
        self.vals = vals

    def get_vals(self):
        return self.vals


def fill_first_system(s, nls, delta=0.1, seed=7):
    """Fill ``s.vals`` with the first Newton system of ``nls`` (same structure as the model ``s``
    was built for) and return the right-hand side."""
    x = nls.x0.copy()
    Fx = np.zeros(nls.nequ)
    nls.residual(x, Fx)
    nls.jac_coord_residual(x, s.Jx_vals)
    cx = np.zeros(nls.ncon)
    nls.cons(x, cx)
    nls.jac_coord(x, s.Jcx_vals)
    lam = np.ones(nls.ncon)
    prepare_newton_system(s, nls, x, lam, Fx, delta)
    rng = np.random.default_rng(seed)
    return rng.standard_normal(nls.nvar + nls.nequ + nls.ncon)


def first_system(nls, method, ctor, delta=0.1, seed=7):
    """Solver + the vals / rhs of the first Newton system (rho = 0, given delta)."""
    s = CaNNOLeSSolver(nls, linsolve=ctor, method=method)
    rhs = fill_first_system(s, nls, delta, seed)
    return s, rhs


def dense_batch_systems(instances, delta=0.1):
    """Config 5: the first Newton systems of the given instances of ``DenseBatchNLS`` (per-instance
    seed 1000 + i).  Returns (structure solver, vals[B, nnz], rhs[B, N]); the COO pattern
    (``s.rows``, ``s.cols``) is shared by all instances."""
    instances = list(instances)
    s = CaNNOLeSSolver(DenseBatchNLS(instances[0] if instances else 0), linsolve=_NoBackend, method="Newton")
    N = s.nvar + s.nequ + s.ncon
    vals = np.empty((len(instances), len(s.vals)))
    rhs = np.empty((len(instances), N))
    for b, i in enumerate(instances):
        rhs[b] = fill_first_system(s, DenseBatchNLS(i), delta, seed=7 + i)
        vals[b] = s.vals
    return s, vals, rhs
