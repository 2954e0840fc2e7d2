"""-m "not gpu": the product's kernel sources (cannoles_b200/csrc/*.cu) compiled for the CPU
emulator in tests/hostsim and checked against the oracle.  This exercises the real symbolic
analysis, launch plan and kernel code paths without a GPU; it is test infrastructure, the
product library itself has no CPU path."""
import os

import numpy as np
import pytest

from cannoles_b200.models import MGH01CON, ExtRosenbrockLinEq, PoissonParamEst
from tests import engine_checks as ec
from tests.problems import EPS, constrained_cases, random_kkt


@pytest.fixture()
def ctor(emu_lib):
    return ec.backend(emu_lib, pin=False)


@pytest.mark.parametrize("ordering", [0, 1, 3])
def test_small_front_path(ctor, oracle_cls, ordering, monkeypatch):
    monkeypatch.setenv("B2_SMALL_MAX_M", "128")
    N, r, c, v = random_kkt(40, 50, 12, 0.1, 21)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 40, 50, 12, ordering=ordering)


@pytest.mark.parametrize("dag", ["1", "0"])
@pytest.mark.parametrize("ordering", [0, 1])
def test_tiled_front_path(ctor, oracle_cls, ordering, monkeypatch, dag):
    monkeypatch.setenv("B2_SMALL_MAX_M", "8")     # force every front of order > 8 onto the tiled path
    monkeypatch.setenv("B2_DAG", dag)             # dataflow kernel (k_front_dag) / k_trsm + k_update launch chain
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    N, r, c, v = random_kkt(50, 60, 15, 0.3, 22)
    B, _ = ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 50, 60, 15, ordering=ordering)
    assert B.stats()["n_large"] > 0


@pytest.mark.parametrize("level_max", ["2", "6", "40"])
def test_cross_level_dataflow_over_the_top_of_the_tree(ctor, oracle_cls, monkeypatch, level_max):
    """The top levels of the assembly tree (those with at most `level_max` fronts) in ONE k_front_dag
    launch -- extend-add tasks, children counters, small fronts cut into tiles like the others --
    below them the per-level launches (small fronts in shared memory, k_trsm / k_update chain)."""
    from cannoles_b200.workloads import first_system
    import functools
    monkeypatch.setenv("B2_SMALL_MAX_M", "24")
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", level_max)
    nls = PoissonParamEst(20)
    s, rhs = first_system(nls, "Newton", functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon))
    B, _ = ec.check_against_oracle(ctor, oracle_cls, s.LDLT.N, s.rows, s.cols, s.vals, nls.nvar, nls.nequ, nls.ncon,
                                   same_perm_tol=1e-8)
    st = B.stats()
    assert st["n_large"] > 0 and st["launches_factor"] < 40


@pytest.mark.parametrize("dag", ["1", "0"])
def test_tiled_path_multi_block_front(ctor, oracle_cls, monkeypatch, dag):
    """One dense front of order 150 on the tiled path: three pivot blocks (64, 64, 22), DMMA tiles;
    with the dataflow kernel: a plain diagonal task, two chain tasks and one ypre task."""
    monkeypatch.setenv("B2_SMALL_MAX_M", "8")
    monkeypatch.setenv("B2_DAG", dag)
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    N, r, c, v = random_kkt(60, 70, 20, 0.5, 71)
    B, _ = ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 60, 70, 20, ordering=1)
    assert B.stats()["max_width"] > 128


@pytest.mark.parametrize("N", [64, 65, 127, 128, 129, 193])
def test_dataflow_kernel_tile_boundaries(ctor, oracle_cls, monkeypatch, N):
    """One dense root front of order N through k_front_dag: full and partial last pivot blocks (N a
    multiple of 64, one more, one less), one to four pivot blocks (plain, chain and ypre tasks)."""
    monkeypatch.setenv("B2_SMALL_MAX_M", "8")
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    nv = N // 3
    ne = N // 2
    nc = N - nv - ne
    Nk, r, c, v = random_kkt(nv, ne, nc, 0.9, 100 + N)
    assert Nk == N
    B, _ = ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, nv, ne, nc, ordering=1)
    assert B.stats()["max_width"] == N


def test_dataflow_kernel_zero_pivot_is_reported_and_does_not_hang(ctor, monkeypatch):
    """An exact zero as the very first pivot of a 193-order front in k_front_dag: the breakdown flag
    comes back, every tile flag is still raised (the tasks behind it run on NaNs instead of waiting)."""
    monkeypatch.setenv("B2_SMALL_MAX_M", "8")
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    nv, ne, nc = 64, 96, 33
    N, r, c, v = random_kkt(nv, ne, nc, 0.9, 293)
    v = v.copy()
    v[(r == 1) & (c == 1)] = 0.0
    B = ctor(N, r, c, v, nvar=nv, nequ=ne, ncon=nc, ordering=1, refine_steps=0)
    assert B.try_to_factorize(v, nv, ne, nc, EPS) is False
    assert B.last_inertia[3] is True


@pytest.mark.parametrize("dag", ["1", "0"])
def test_tiled_path_multi_chunk_trsm(ctor, oracle_cls, monkeypatch, dag):
    """Order-260 dense front: the rows below the first pivot block span two k_trsm CTAs, so the
    factored diagonal block must come from the staging area, not from the panel being read
    (launch chain); five row blocks with a contribution block (dataflow kernel)."""
    monkeypatch.setenv("B2_SMALL_MAX_M", "8")
    monkeypatch.setenv("B2_DAG", dag)
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    N, r, c, v = random_kkt(100, 130, 30, 0.5, 73)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 100, 130, 30, ordering=1)


def test_adaptive_refinement_stops_early(ctor):
    N, r, c, v = random_kkt(30, 40, 8, 0.2, 74)
    B = ctor(N, r, c, v, nvar=30, nequ=40, ncon=8, refine_steps=3)
    assert B.try_to_factorize(v, 30, 40, 8, EPS)
    d = np.zeros(N)
    B.solve_ldl(np.ones(N), d)
    assert B.last_relres <= 1e-13 and B.last_sweeps <= 2
    C = ctor(N, r, c, v, nvar=30, nequ=40, ncon=8, refine_steps=2, refine_tol=0.0)   # fixed sweeps
    assert C.try_to_factorize(v, 30, 40, 8, EPS)
    C.solve_ldl(np.ones(N), d)
    assert C.last_sweeps == 3


@pytest.mark.parametrize("case", [(40, 50, 12, 0.1, 21, 0), (60, 70, 20, 0.5, 71, 1), (45, 60, 10, 0.8, 72, 0)])
def test_multi_cta_solve_of_big_fronts(ctor, oracle_cls, monkeypatch, case):
    """Every front of order > 4 through k_fwd_big / k_bwd_big (flag-chained 64-row chunks): fronts
    with children, rows below the pivots, and a 150-wide pivot block spanning three chunks."""
    monkeypatch.setenv("B2_SOLVE_BIG_M", "4")
    nv, ne, nc, dens, seed, order = case
    N, r, c, v = random_kkt(nv, ne, nc, dens, seed)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, nv, ne, nc, ordering=order)


def test_lookahead_diagonal_factorization_path(ctor, oracle_cls, monkeypatch):
    """The optional look-ahead schedule (k_diag on a side stream, prefactored k_trsm, k_update that
    leaves the next diagonal block alone) gives the same factorization."""
    monkeypatch.setenv("B2_SMALL_MAX_M", "8")
    monkeypatch.setenv("B2_DAG", "0")
    monkeypatch.setenv("B2_LOOKAHEAD", "1")
    N, r, c, v = random_kkt(60, 70, 20, 0.5, 71)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 60, 70, 20, ordering=1)
    N, r, c, v = random_kkt(50, 60, 15, 0.3, 22)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 50, 60, 15, ordering=0)


def test_small_delta_refinement_reaches_the_bar(ctor, oracle_cls):
    """delta = sqrt(eps): refinement sweeps are taken until the relative residual is <= 1e-12."""
    N, r, c, v = random_kkt(40, 50, 14, 0.1, 79, delta=1.4901161193847656e-08)
    B = ctor(N, r, c, v, nvar=40, nequ=50, ncon=14, refine_steps=3)
    assert B.try_to_factorize(v, 40, 50, 14, EPS)
    O = oracle_cls(N, r, c, v, perm=B.perm)
    assert O.try_to_factorize(v, 40, 50, 14, EPS)
    rhs = np.random.default_rng(8).standard_normal(N)
    d = np.zeros(N)
    B.solve_ldl(rhs, d)
    assert B.last_relres <= ec.RESID_TOL
    assert np.linalg.norm(O.matvec(d) + rhs) <= ec.RESID_TOL * np.linalg.norm(rhs)


def test_golden_vectors(ctor):
    for name in ("mgh01con_first_kkt", "random_kkt_0", "random_kkt_1"):
        ec.check_golden(ctor, name)


def test_zero_pivot_is_reported_not_raised(ctor, oracle_cls):
    """MGH01CON's first KKT matrix in natural order: d = [88, 0, ...] -> False, never an exception."""
    rows = np.array([1, 1, 2, 2, 3, 4, 4, 5, 3, 4, 5, 1, 2], dtype=np.int64)
    cols = np.array([1, 1, 1, 2, 1, 1, 2, 1, 3, 4, 5, 1, 2], dtype=np.int64)
    vals = np.array([88., 0, 0, 0, -1, 24, 10, 1, -1, -1, -0.1, 0, 0])
    B = ctor(5, rows, cols, vals, nvar=2, nequ=2, ncon=1, ordering=1)
    assert B.try_to_factorize(vals, 2, 2, 1, EPS) is False
    assert B.last_inertia[3] is True and B.last_inertia[1] >= 1
    vals[-2:] = EPS ** (1 / 3)                    # the rho_0 retry of newton_system!
    assert B.try_to_factorize(vals, 2, 2, 1, EPS) is True
    assert B.n_shift == 1


def test_retry_speculation_is_validated(ctor, oracle_cls):
    N, r, c, v = random_kkt(30, 35, 9, 0.2, 27)
    ec.check_retry_is_validated(ctor, oracle_cls, N, r, c, v, 30, 35, 9, rho=6.0554544523933395e-06)


def test_shift_retry_is_bit_identical(ctor, oracle_cls):
    N, r, c, v = random_kkt(30, 35, 9, 0.15, 23)
    ec.check_shift_path(ctor, oracle_cls, N, r, c, v, 30, 35, 9, rho=6.0554544523933395e-06)


def test_malformed_input_rejected(ctor):
    from cannoles_b200.linsolve import B200Error
    rows = np.array([1, 1], dtype=np.int64)
    cols = np.array([1, 2], dtype=np.int64)      # strictly upper entry
    with pytest.raises(B200Error, match="upper"):
        ctor(2, rows, cols, np.ones(2), nvar=2, nequ=0, ncon=0)
    with pytest.raises(B200Error, match="range"):
        ctor(2, np.array([3, 1], dtype=np.int64), np.array([1, 1], dtype=np.int64), np.ones(2), nvar=2, nequ=0, ncon=0)


def test_config_slices(ctor, oracle_cls):
    """Small instances of configs C2 and C4 through the first Newton system."""
    from scripts.gpu_check import first_system
    for nls, method in ((ExtRosenbrockLinEq(60), "Newton_noFHess"), (PoissonParamEst(6), "Newton")):
        import functools
        s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon))
        ec.check_against_oracle(ctor, oracle_cls, s.LDLT.N, s.rows, s.cols, s.vals, nls.nvar, nls.nequ, nls.ncon)


def test_cannoles_loop_same_iterations_as_oracle(ctor):
    """Iteration count, nfact, nlinsolve and the final point equal the oracle's on the same order."""
    nls = MGH01CON()
    stb, sto = ec.run_cannoles_both(nls, ctor)
    ec.assert_same_run(stb, sto)
    for nls, xf in constrained_cases()[:2]:
        stb, sto = ec.run_cannoles_both(nls, ctor)
        ec.assert_same_run(stb, sto)
        assert np.allclose(stb.solution, xf, atol=1e-4)
