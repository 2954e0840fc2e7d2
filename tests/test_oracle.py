"""The oracle pinned against everything the reference offers for this path:
hand-derived first KKT system of MGH01CON (App. D), dense-LAPACK cross-checks, and the
end-to-end known answers of reference/test/runtests.jl through the restated `cannoles` loop."""
import numpy as np
import pytest

from cannoles_b200 import CaNNOLeSSolver, cannoles, register_linsolve, solve
from cannoles_b200.models import MGH01CON, MGH01_noFHess
from tests.problems import (EPS, constrained_cases, dense_from_coo, hs6, random_kkt,
                            unconstrained_cases)

MGH_ROWS = np.array([1, 1, 2, 2, 3, 4, 4, 5, 3, 4, 5, 1, 2], dtype=np.int64)
MGH_COLS = np.array([1, 1, 1, 2, 1, 1, 2, 1, 3, 4, 5, 1, 2], dtype=np.int64)
MGH_VALS = np.array([88., 0, 0, 0, -1, 24, 10, 1, -1, -1, -0.1, 0, 0])


@pytest.fixture(autouse=True)
def _register(oracle_cls):
    register_linsolve("ldlfactorizations", oracle_cls)


def test_mgh01con_first_kkt_system(oracle_cls):
    """order (lambda, r1, x1, r2, x2) gives D = [-0.1, -1, 99, -6.8181.., 14.666..]
    and d = -K^-1 [dual; primal]."""
    vals = MGH_VALS.copy()
    L = oracle_cls(5, MGH_ROWS, MGH_COLS, vals, perm=[4, 2, 0, 3, 1])
    assert L.nnzA == 10
    assert L.try_to_factorize(vals, 2, 2, 1, EPS)
    np.testing.assert_allclose(L.factor.d, [-0.1, -1, 99, -75 / 11, 44 / 3], rtol=1e-14)
    assert L.inertia(EPS) == (2, 0, 3)
    d = np.zeros(5)
    assert L.solve_ldl(np.array([0, -44, 0, 0, -1.2]), d)
    np.testing.assert_allclose(d, [-52 / 55, 149 / 55, 52 / 55, 4.4, -236 / 11], rtol=1e-13)


def test_mgh01con_breakdown_orders(oracle_cls):
    """K22 == 0 exactly: any order eliminating x2 before r2 meets a zero pivot at rho = 0."""
    vals = MGH_VALS.copy()
    L = oracle_cls(5, MGH_ROWS, MGH_COLS, vals, perm=[0, 1, 2, 3, 4])
    assert not L.try_to_factorize(vals, 2, 2, 1, EPS)
    assert L.factor.d[0] == 88.0 and L.factor.d[1] == 0.0


def test_set_vals_variants_agree(oracle_cls):
    N, r, c, v = random_kkt(30, 40, 10, 0.2, 3)
    A = oracle_cls(N, r, c, v)
    B = oracle_cls(N, r, c, v, use_search=True)
    A.set_vals(v)
    B.set_vals(v)
    assert np.array_equal(A.nzval, B.nzval)
    K = dense_from_coo(N, r, c, v)
    cp, rv, nz = A.colptr, A.rowval, A.nzval
    for j in range(N):
        for p in range(cp[j], cp[j + 1]):
            assert abs(K[rv[p], j] - nz[p]) <= 1e-15 * max(1, abs(nz[p]))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_inertia_and_solve_against_lapack(oracle_cls, seed):
    nv, ne, nc = 25 + seed, 30, 7
    N, r, c, v = random_kkt(nv, ne, nc, 0.15, seed)
    K = dense_from_coo(N, r, c, v)
    L = oracle_cls(N, r, c, v)
    assert L.try_to_factorize(v, nv, ne, nc, EPS)
    ev = np.linalg.eigvalsh(K)
    assert L.inertia(EPS) == (int((ev > 0).sum()), 0, int((ev < 0).sum()))
    rhs = np.random.default_rng(seed).standard_normal(N)
    d = np.zeros(N)
    L.solve_ldl(rhs, d)
    np.testing.assert_allclose(d, -np.linalg.solve(K, rhs), rtol=1e-10, atol=1e-12)
    assert np.linalg.norm(L.matvec(d) + rhs) <= 1e-12 * np.linalg.norm(rhs)


def test_amd_is_a_permutation_and_reduces_fill(oracle_cls):
    import scipy.sparse as sp
    n = 24
    G = sp.diags([1., 1.], [1, -1], shape=(n, n))
    A = sp.tril(-(sp.kron(G, sp.identity(n)) + sp.kron(sp.identity(n), G)) + 4 * sp.identity(n * n)).tocoo()
    v = A.data.copy()
    amd = oracle_cls(n * n, A.row + 1, A.col + 1, v)
    nat = oracle_cls(n * n, A.row + 1, A.col + 1, v, ordering=1)
    assert sorted(amd.perm.tolist()) == list(range(n * n))
    assert amd.nnzL < 0.5 * nat.nnzL
    assert amd.try_to_factorize(v, n * n, 0, 0, EPS) and nat.try_to_factorize(v, n * n, 0, 0, EPS)


# ---- end-to-end known answers of the reference's own tests -----------------------------------

def test_reference_unconstrained_known_answers():
    for nls, xf in unconstrained_cases():
        st = cannoles(nls, linsolve="ldlfactorizations")
        assert np.allclose(st.solution, xf, atol=1e-4), (nls.name, st.solution)


def test_reference_constrained_known_answers():
    for nls, xf in constrained_cases():
        st = cannoles(nls, linsolve="ldlfactorizations")
        assert np.allclose(st.solution, xf, atol=1e-4), st.solution


def test_reference_resolve_and_small_residual():
    nls = hs6()
    solver = CaNNOLeSSolver(nls, linsolve="ldlfactorizations")
    st = solve(solver, nls)
    assert st.status == "first_order" and np.allclose(st.solution, [1, 1], atol=1e-6)
    st = solve(solver, nls, x=np.array([10.0, 10.0]))
    assert st.status == "first_order" and np.allclose(st.solution, [1, 1], atol=1e-6)
    st = solve(solver, nls, atol=1e-15, rtol=0.0, Fatol=1e-6, Frtol=0.0)
    assert st.status == "small_residual" and abs(st.objective) < 1e-6
    st = solve(solver, nls, x=np.array([0.99999, 0.99999]), atol=1e-15, rtol=0.0, Fatol=1e-6, Frtol=0.0)
    assert st.status == "small_residual" and abs(st.objective) < 1e-6
    nlp = hs6(shift=True)   # same sparsity: reset!(solver, nlp) keeps the analysed pattern
    st = solve(solver, nlp)
    assert st.status == "first_order" and np.allclose(st.solution, [0, 0], atol=1e-6)


def test_reference_mgh01con_and_gauss_newton():
    nls = MGH01CON()
    st = cannoles(nls, linsolve="ldlfactorizations")
    assert st.status == "first_order"
    assert abs(st.solution[0]) < 1e-8          # the constraint is x1 == 0 (lcon = 0)
    nls.x0[:] = 0.0                             # test/runtests.jl:28: no factorization at all
    st = cannoles(nls, linsolve="ldlfactorizations")
    assert st.status == "first_order" and st.solver_specific["nfact"] == 0
    g = MGH01_noFHess()
    st = cannoles(g, linsolve="ldlfactorizations", method="Newton_noFHess")
    assert np.allclose(st.solution, [1, 1], atol=1e-6)
    with pytest.raises(TypeError):
        cannoles(MGH01_noFHess(), linsolve="ldlfactorizations")   # default :Newton -> MethodError


def test_reference_api_errors_and_callback():
    with pytest.raises(ValueError, match="`method` must be one of these"):
        cannoles(hs6(), linsolve="ldlfactorizations", method="truc")
    nls = hs6()
    nls.minimize = False
    with pytest.raises(ValueError, match="only works for minimization"):
        cannoles(nls, linsolve="ldlfactorizations")
    from cannoles_b200.models import SymbolicNLSModel
    nls = SymbolicNLSModel(lambda x: [x[0] - 1, 10 * (x[1] - x[0] ** 2)], [-1.2, 1.0], lambda x: [x[0] * x[1] - 1])

    def cb(nls_, solver_, stats):
        if stats.iter == 4:
            stats.status = "user"
    st = cannoles(nls, linsolve="ldlfactorizations", callback=cb)
    assert st.iter == 4 and st.status == "user"
