"""Parity checks shared by the CPU-emulator tests (-m "not gpu") and the GPU tests (-m gpu):
the B200 backend against the oracle on the same inputs, through the reference's verbs."""
import functools
import os

import numpy as np

from cannoles_b200.linsolve import B200Struct
from tests.problems import EPS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RESID_TOL = 1e-12     # north_star: KKT solve relative residual <= 1e-12


def backend(lib=None, **kw):
    return functools.partial(B200Struct, _lib=lib, **kw) if lib is not None else functools.partial(B200Struct, **kw)


def check_against_oracle(ctor, oracle_cls, N, rows, cols, vals, nvar, nequ, ncon, ordering=0,
                         same_perm_tol=1e-9, expect_ok=True):
    B = ctor(N, rows, cols, vals, nvar=nvar, nequ=nequ, ncon=ncon, ordering=ordering, refine_steps=1)
    ok = B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    O = oracle_cls(N, rows, cols, vals, perm=B.perm)      # same elimination order
    ok_o = O.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert ok == ok_o == expect_ok
    # kernel (1): CSC values bit-exact, same pattern
    cp, rv = B.csc()
    assert np.array_equal(cp, O.colptr) and np.array_equal(rv, O.rowval)
    assert np.array_equal(B.nzval, O.nzval)
    # symbolic: nnz(L) of the same elimination order (integer work: equal)
    assert B.stats()["nnzL"] == O.nnzL
    # inertia bit-exact
    assert B.last_inertia[:3] == O.inertia(EPS)
    if not ok:
        return B, O
    dB, dO = B.factor.d, O.factor.d
    assert np.max(np.abs(dB - dO) / np.abs(dO)) < same_perm_tol
    rhs = np.random.default_rng(5).standard_normal(N)
    xb, xo = np.zeros(N), np.zeros(N)
    assert B.solve_ldl(rhs, xb) and O.solve_ldl(rhs, xo)
    assert B.last_relres <= RESID_TOL
    assert np.linalg.norm(O.matvec(xb) + rhs) <= RESID_TOL * np.linalg.norm(rhs)   # d = -K^-1 rhs
    assert np.linalg.norm(xb - xo) <= same_perm_tol * np.linalg.norm(xo)
    return B, O


def check_golden(ctor, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    nvar, nequ, ncon = (int(t) for t in z["dims"])
    vals = z["vals"].copy()
    B = ctor(int(z["N"]), z["rows"], z["cols"], vals, nvar=nvar, nequ=nequ, ncon=ncon, perm=z["perm"])
    ok = B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert ok == bool(z["ok"])
    assert np.array_equal(B.nzval, z["nzval"])
    assert tuple(B.last_inertia[:3]) == tuple(int(t) for t in z["inertia"])
    # the backend postorders the supplied order (an equivalent reordering): compare the pivot
    # attached to each ORIGINAL variable
    dB, dG = np.empty(int(z["N"])), np.empty(int(z["N"]))
    dB[B.perm] = B.factor.d
    dG[z["perm"]] = z["D"]
    np.testing.assert_allclose(dB, dG, rtol=1e-10)
    d = np.zeros(int(z["N"]))
    B.solve_ldl(z["rhs"].copy(), d)
    np.testing.assert_allclose(d, z["d"], rtol=1e-9, atol=1e-12)


def check_shift_path(ctor, oracle_cls, N, rows, cols, vals, nvar, nequ, ncon, rho, ordering=0):
    """A rho retry through b2_refactorize_shift must equal a full re-upload bit for bit."""
    vals = vals.copy()
    B = ctor(N, rows, cols, vals, nvar=nvar, nequ=nequ, ncon=ncon, ordering=ordering)
    B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    vals[-nvar:] = rho
    B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert B.n_shift == 1 and B.n_upload == 1
    nz_shift, d_shift, in_shift = B.nzval, B.factor.d, B.last_inertia
    C = ctor(N, rows, cols, vals, nvar=nvar, nequ=nequ, ncon=ncon, ordering=ordering, shift_retries=False)
    C.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert C.n_shift == 0
    assert np.array_equal(nz_shift, C.nzval)
    assert np.array_equal(d_shift, C.factor.d)
    assert in_shift == C.last_inertia
    O = oracle_cls(N, rows, cols, vals, perm=B.perm)
    O.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert np.array_equal(nz_shift, O.nzval)


def check_retry_is_validated(ctor, oracle_cls, N, rows, cols, vals, nvar, nequ, ncon, rho, ordering=0):
    """A caller that does NOT follow newton_system!'s protocol: it sets the rho segment to a non-zero
    constant (looks like a retry) but also edits H / J.  The speculative shift must be discarded and
    the result must be the factorization of what was passed, bit for bit."""
    vals = vals.copy()
    B = ctor(N, rows, cols, vals, nvar=nvar, nequ=nequ, ncon=ncon, ordering=ordering)
    B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    vals[-nvar:] = rho
    vals[: len(vals) // 3] *= 1.25                     # LM-style caller: other segments change too
    B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert B.n_shift == 0 and B.n_upload == 2          # speculation rejected, full factorization returned
    C = ctor(N, rows, cols, vals, nvar=nvar, nequ=nequ, ncon=ncon, ordering=ordering, shift_retries=False)
    C.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert np.array_equal(B.nzval, C.nzval) and np.array_equal(B.factor.d, C.factor.d)
    assert B.last_inertia == C.last_inertia
    O = oracle_cls(N, rows, cols, vals, perm=B.perm)
    O.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert np.array_equal(B.nzval, O.nzval)
    # a non-constant trailing segment whose ends happen to agree is not a retry either
    vals[-nvar:] = rho * 2
    vals[-nvar + 1] = rho * 3
    B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    C.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert np.array_equal(B.nzval, C.nzval) and np.array_equal(B.factor.d, C.factor.d)
    # and a genuine retry after all that is taken by the shift path again
    vals[-nvar:] = rho * 4
    B.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    C.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    assert B.n_shift == 1
    assert np.array_equal(B.nzval, C.nzval) and np.array_equal(B.factor.d, C.factor.d)


def run_cannoles_both(nls, ctor, method="Newton", **kw):
    """The restated cannoles loop with the B200 backend and with the oracle on the same order."""
    from cannoles_b200 import CaNNOLeSSolver, solve
    from oracle import LDLFactStruct
    sb = CaNNOLeSSolver(nls, linsolve=functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon),
                        method=method)
    perm = sb.LDLT.perm
    stb = solve(sb, nls, **kw)
    nls.reset_counters()
    so = CaNNOLeSSolver(nls, linsolve=functools.partial(LDLFactStruct, perm=perm), method=method)
    sto = solve(so, nls, **kw)
    nls.reset_counters()
    return stb, sto


def assert_same_run(stb, sto, rtol=1e-8):
    """north_star: iteration count equal; nfact / nlinsolve equal; x, objective, ||c|| within 1e-8."""
    assert stb.status == sto.status
    assert stb.iter == sto.iter
    assert stb.solver_specific["nfact"] == sto.solver_specific["nfact"]
    assert stb.solver_specific["nlinsolve"] == sto.solver_specific["nlinsolve"]
    scale = max(1.0, float(np.linalg.norm(sto.solution)))
    assert np.linalg.norm(stb.solution - sto.solution) <= rtol * scale
    assert abs(stb.objective - sto.objective) <= rtol * max(1.0, abs(sto.objective))
    assert abs(stb.primal_feas - sto.primal_feas) <= rtol * max(1.0, abs(sto.primal_feas))
