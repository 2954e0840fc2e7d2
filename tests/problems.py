"""Shared problem builders for the tests: the reference's own test problems
(reference/test/runtests.jl:57-100) and small random KKT systems."""
import numpy as np
import scipy.sparse as sp

from cannoles_b200.models import SymbolicNLSModel

EPS = 2.0 ** -52


def F_linear(x): return [x[0] - 2, x[1] - 3]
def F_Rosen(x): return [x[0] - 1, 10 * (x[1] - x[0] ** 2)]
def F_larger(x, n): return [10 * (x[i + 1] - x[i] ** 2) for i in range(n - 1)] + [x[i] - 1 for i in range(n - 1)]
def F_under(x, n): return [x[0] - x[i] for i in range(1, n)]
def c_linear(x): return [sum(x) - 1]


def c_quad(x):
    p = 1
    for xi in x:
        p = p * xi
    return [sum(xi ** 2 for xi in x) - 5, p - 2]


def unconstrained_cases(n=10):
    """test/runtests.jl:65-72: (model, known solution)."""
    cases = [(SymbolicNLSModel(F_linear, -np.ones(2)), [2.0, 3.0]),
             (SymbolicNLSModel(F_Rosen, [-1.2, 1.0]), [1.0, 1.0]),
             (SymbolicNLSModel(lambda x: F_larger(x, n), 0.9 * np.ones(n)), np.ones(n))]
    for i in range(1, 6):
        cases.append((SymbolicNLSModel(lambda x: F_under(x, n), i * np.ones(n)), i * np.ones(n)))
    return cases


def constrained_cases(n=10):
    """test/runtests.jl:82-91."""
    return [
        (SymbolicNLSModel(F_linear, -np.ones(2), c_linear), [0.0, 1.0]),
        (SymbolicNLSModel(F_Rosen, [-1.2, 1.0], c_linear), [0.6188, 0.3812]),
        (SymbolicNLSModel(lambda x: F_under(x, n), np.arange(1, n + 1) / n, c_linear), np.ones(n) / n),
        (SymbolicNLSModel(F_linear, [0.9, 1.9], c_quad), [1.0, 2.0]),
        (SymbolicNLSModel(F_Rosen, [0.9, 1.9], c_quad), [1.0, 2.0]),
        (SymbolicNLSModel(lambda x: F_larger(x, 3), [0.5, 1.0, 1.5], c_quad), [1.0647, 1.215, 1.546]),
    ]


def hs6(shift=False):
    """test/runtests.jl:116-125 ('HS6') and :187-195 ('shifted HS6')."""
    F = (lambda x: [x[0]]) if shift else (lambda x: [x[0] - 1])
    return SymbolicNLSModel(F, [-1.2, 1.0], lambda x: [10 * (x[1] - x[0] ** 2)])


def random_kkt(nv, ne, nc, dens, seed, delta=0.1, hscale=0.1):
    """COO lower triangle (1-based) of a quasi-definite [H+2I, Jx', Jc'; Jx, -I, 0; Jc, 0, -delta I]
    in the segment order of src/CaNNOLeS.jl:281-315 (with duplicates on the (1,1) diagonal)."""
    rng = np.random.default_rng(seed)
    H = sp.random(nv, nv, dens, random_state=seed)
    H = sp.tril((H + H.T) * hscale + 2 * sp.identity(nv)).tocoo()
    Jx = (sp.random(ne, nv, dens, random_state=seed + 1) + sp.random(ne, nv, min(1.0, 2.0 / nv), random_state=seed + 3)).tocoo()
    Jc = sp.random(nc, nv, max(dens, min(1.0, 2.0 / nv)), random_state=seed + 2).tocoo()
    rows = np.concatenate([H.row, Jx.row + nv, Jc.row + nv + ne, nv + np.arange(ne), nv + ne + np.arange(nc), np.arange(nv)]) + 1
    cols = np.concatenate([H.col, Jx.col, Jc.col, nv + np.arange(ne), nv + ne + np.arange(nc), np.arange(nv)]) + 1
    vals = np.concatenate([H.data, Jx.data, Jc.data, -np.ones(ne), -delta * np.ones(nc), np.zeros(nv)])
    del rng
    return nv + ne + nc, rows.astype(np.int64), cols.astype(np.int64), vals.astype(np.float64)


def dense_from_coo(N, rows, cols, vals):
    K = np.zeros((N, N))
    for r, c, v in zip(rows - 1, cols - 1, vals):
        K[r, c] += v
        if r != c:
            K[c, r] += v
    return K
