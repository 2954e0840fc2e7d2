"""-m gpu: the CUDA path (libcannoles_b200.so through the C ABI) against the oracle on the same
seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties (expected inertia, residual, determinism, linearity)."""
import functools

import numpy as np
import pytest

from cannoles_b200.models import MGH01CON, ExtRosenbrockLinEq, PoissonParamEst
from tests import engine_checks as ec
from tests.problems import EPS, constrained_cases, hs6, random_kkt, unconstrained_cases

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctor(gpu_lib):
    assert gpu_lib.b2_device_count() > 0
    return ec.backend()


def test_library_is_the_cuda_build(gpu_lib):
    from cannoles_b200 import _capi
    assert _capi.LIB_PATH.endswith("csrc/libcannoles_b200.so")
    for name in _capi.EXPORTS:
        assert hasattr(gpu_lib, name), name


def test_golden_vectors(ctor):
    for name in ("mgh01con_first_kkt", "random_kkt_0", "random_kkt_1", "random_kkt_2"):
        ec.check_golden(ctor, name)


@pytest.mark.parametrize("ordering", [0, 1, 3])
@pytest.mark.parametrize("seed", [31, 32])
def test_random_kkt_against_oracle(ctor, oracle_cls, ordering, seed, monkeypatch):
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")     # every tiled front through k_front_dag, one-block fronts too
    nv, ne, nc = 300, 400, 80
    N, r, c, v = random_kkt(nv, ne, nc, 0.02, seed)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, nv, ne, nc, ordering=ordering)


@pytest.mark.parametrize("dag", ["1", "0"])
def test_dense_front_tiled_path_against_oracle(ctor, oracle_cls, monkeypatch, dag):
    """One dense 700-order front: five+ pivot blocks, 11 x 11 tiles; through the dataflow kernel
    (k_front_dag: chain / ypre / plain tile tasks) and through the k_trsm / k_update launch chain."""
    monkeypatch.setenv("B2_DAG", dag)
    nv, ne, nc = 250, 350, 100
    N, r, c, v = random_kkt(nv, ne, nc, 0.6, 33)
    B, _ = ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, nv, ne, nc, ordering=1)
    assert B.stats()["n_large"] >= 1 and B.stats()["max_front"] >= 600


@pytest.mark.parametrize("N", [64, 65, 127, 128, 129, 192, 193, 257, 320])
def test_dataflow_kernel_tile_boundaries(ctor, oracle_cls, monkeypatch, N):
    """One dense root front of order N through k_front_dag: full and partial last pivot blocks,
    one to five pivot blocks (plain, chain and ypre tasks)."""
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    nv = N // 3
    ne = N // 2
    nc = N - nv - ne
    Nk, r, c, v = random_kkt(nv, ne, nc, 0.9, 100 + N)
    assert Nk == N
    B, _ = ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, nv, ne, nc, ordering=1)
    assert B.stats()["max_width"] == N


@pytest.mark.timeout(120)
def test_dataflow_kernel_zero_pivot_is_reported_and_does_not_hang(ctor, monkeypatch):
    """An exact zero as the very first pivot of a 193-order front in k_front_dag: the breakdown flag
    comes back, every tile flag is still raised (the tasks behind it run on NaNs instead of waiting)."""
    monkeypatch.setenv("B2_DAG_LEVEL_MAX", "1000000000")
    nv, ne, nc = 64, 96, 33
    N, r, c, v = random_kkt(nv, ne, nc, 0.9, 293)
    v = v.copy()
    v[(r == 1) & (c == 1)] = 0.0
    B = ctor(N, r, c, v, nvar=nv, nequ=ne, ncon=nc, ordering=1, refine_steps=0)
    assert B.try_to_factorize(v, nv, ne, nc, EPS) is False
    assert B.last_inertia[3] is True


def test_empty_blocks_and_tiny_systems(ctor, oracle_cls):
    """ncon = 0 (unconstrained: no delta segment) and a 1 x 1 system."""
    N, r, c, v = random_kkt(12, 20, 0, 0.3, 34)
    ec.check_against_oracle(ctor, oracle_cls, N, r, c, v, 12, 20, 0)
    one = np.array([1], dtype=np.int64)
    ec.check_against_oracle(ctor, oracle_cls, 1, one, one, np.array([2.0]), 1, 0, 0)


def test_zero_pivot_and_wrong_inertia_return_false(ctor, oracle_cls):
    rows = np.array([1, 1, 2, 2, 3, 4, 4, 5, 3, 4, 5, 1, 2], dtype=np.int64)
    cols = np.array([1, 1, 1, 2, 1, 1, 2, 1, 3, 4, 5, 1, 2], dtype=np.int64)
    vals = np.array([88., 0, 0, 0, -1, 24, 10, 1, -1, -1, -0.1, 0, 0])
    B = ctor(5, rows, cols, vals, nvar=2, nequ=2, ncon=1, ordering=1)
    assert B.try_to_factorize(vals, 2, 2, 1, EPS) is False and B.last_inertia[3] is True
    # indefinite (1,1) block, no zero pivot: inertia (1, 0, 1) instead of (2, 0, 0)
    r2 = np.array([1, 2, 1, 2], dtype=np.int64)
    v2 = np.array([1.0, -3.0, 0.0, 0.0])
    C = ctor(2, r2, r2, v2, nvar=2, nequ=0, ncon=0)
    assert C.try_to_factorize(v2, 2, 0, 0, EPS) is False
    assert C.last_inertia == (1, 0, 1, False)
    v2[2:] = 5.0                                   # rho = 5 makes it positive definite
    assert C.try_to_factorize(v2, 2, 0, 0, EPS) is True and C.n_shift == 1


def test_shift_retry_is_bit_identical(ctor, oracle_cls):
    N, r, c, v = random_kkt(200, 260, 50, 0.03, 35)
    ec.check_shift_path(ctor, oracle_cls, N, r, c, v, 200, 260, 50, rho=6.0554544523933395e-06)


def test_retry_speculation_is_validated_on_the_device(ctor, oracle_cls):
    N, r, c, v = random_kkt(200, 260, 50, 0.03, 36)
    ec.check_retry_is_validated(ctor, oracle_cls, N, r, c, v, 200, 260, 50, rho=6.0554544523933395e-06)


def test_config_slices_against_oracle(ctor, oracle_cls):
    from scripts.gpu_check import first_system
    for nls, method in ((ExtRosenbrockLinEq(20_000), "Newton_noFHess"), (PoissonParamEst(96), "Newton")):
        s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon))
        ec.check_against_oracle(ctor, oracle_cls, s.LDLT.N, s.rows, s.cols, s.vals, nls.nvar, nls.nequ,
                                nls.ncon, same_perm_tol=1e-8)


def test_reference_known_answers_through_the_gpu_backend(ctor):
    """reference/test/runtests.jl:65-100 with linsolve = b200; and the same iteration counts,
    nfact, nlinsolve and final point as the oracle on the same elimination order."""
    for nls, xf in unconstrained_cases() + constrained_cases():
        stb, sto = ec.run_cannoles_both(nls, ctor)
        assert np.allclose(stb.solution, xf, atol=1e-4)
        ec.assert_same_run(stb, sto)
    stb, sto = ec.run_cannoles_both(MGH01CON(), ctor)
    ec.assert_same_run(stb, sto)
    # HS6 (test/runtests.jl:116-138, atol 1e-6 there): the last Newton systems have delta = 1e-5 and the
    # un-refined reference solve (LDLFactorizations does no refinement) leaves an error of ~2e-8 in x under the
    # elimination order in use; the backend's adaptive refinement lands on [1, 1] to the last bit.  Same
    # iteration counts, and the two final points agree within the reference's own solve error
    stb, sto = ec.run_cannoles_both(hs6(), ctor)
    assert stb.status == "first_order" and np.allclose(stb.solution, [1, 1], atol=1e-6)
    ec.assert_same_run(stb, sto, rtol=1e-7)
    assert np.linalg.norm(stb.solution - 1.0) <= np.linalg.norm(sto.solution - 1.0) + 1e-12
    # without refinement the backend does the reference's arithmetic: 1e-8 holds
    stb0, sto0 = ec.run_cannoles_both(hs6(), functools.partial(ctor, refine_steps=0))
    ec.assert_same_run(stb0, sto0, rtol=1e-8)


def test_cannoles_on_a_config_slice_matches_oracle(ctor):
    nls = PoissonParamEst(24)
    stb, sto = ec.run_cannoles_both(nls, ctor, max_time=600.0)
    ec.assert_same_run(stb, sto)
    nls = ExtRosenbrockLinEq(2000)
    stb, sto = ec.run_cannoles_both(nls, ctor, method="Newton_noFHess", max_time=600.0)
    ec.assert_same_run(stb, sto)


@pytest.mark.parametrize("cfg", ["c2", "c4"])
def test_full_size_properties_and_oracle(ctor, oracle_cls, cfg):
    """BASELINE.json full sizes with the BENCHED ordering (nested dissection): expected inertia
    (nvar, 0, nequ+ncon), residual <= 1e-12, run-to-run determinism of the pivots, linearity of the
    solve -- and the oracle run on the same elimination order: nnz(L) equal, inertia equal, pivots
    within 1e-9, step within 1e-8 (reference/src/solver_types.jl:89-96, :69-77)."""
    from scripts.gpu_check import first_system
    nls, method = ((ExtRosenbrockLinEq(100_000), "Newton_noFHess") if cfg == "c2"
                   else (PoissonParamEst(512), "Newton"))
    s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                                                         ordering=0, shift_retries=False))
    B, N = s.LDLT, s.LDLT.N
    assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    assert B.last_inertia == (nls.nvar, 0, nls.nequ + nls.ncon, False)
    d1 = B.factor.d
    assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    assert np.array_equal(d1, B.factor.d)                       # deterministic, no atomics on values
    x1, x2, x3 = np.zeros(N), np.zeros(N), np.zeros(N)
    b2 = np.random.default_rng(9).standard_normal(N)
    B.solve_ldl(rhs, x1)
    assert B.last_relres <= ec.RESID_TOL
    B.solve_ldl(b2, x2)
    B.solve_ldl(2.5 * rhs + b2, x3)
    assert np.linalg.norm(x3 - (2.5 * x1 + x2)) <= 1e-9 * np.linalg.norm(x3)
    # the oracle on the same elimination order
    O = oracle_cls(N, s.rows, s.cols, s.vals, perm=B.perm)
    assert O.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    assert B.stats()["nnzL"] == O.nnzL
    assert np.array_equal(B.nzval, O.nzval)
    assert B.last_inertia[:3] == O.inertia(EPS)
    dO = O.factor.d
    assert np.max(np.abs(d1 - dO) / np.abs(dO)) < 1e-9
    xo = np.zeros(N)
    O.solve_ldl(rhs, xo)
    assert np.linalg.norm(x1 - xo) <= 1e-8 * np.linalg.norm(xo)
    assert np.linalg.norm(O.matvec(x1) + rhs) <= ec.RESID_TOL * np.linalg.norm(rhs)


def _rho_retry(B, s, nls):
    """The caller protocol of newton_system! (reference/src/CaNNOLeS.jl:1023-1043): rho = 0, then rho0,
    then x 100 while the inertia is wrong.  Returns (ok, rho, tries)."""
    ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    rho, tries = 0.0, 0
    while not ok and tries < 6:
        rho = EPS ** (1.0 / 3.0) if rho == 0.0 else 100.0 * rho
        s.vals[len(s.vals) - nls.nvar:] = rho
        ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
        tries += 1
    return ok, rho, tries


@pytest.mark.timeout(900)
def test_c3_full_size_properties(ctor):
    """BASELINE config 3 at its full size (50 k cameras, 1 M points, N = 7 450 007, Gauss-Newton: zero
    (1,1) block, so the rho = 0 attempt breaks down and the reference's rho0 retry is taken):
    expected inertia, residual <= 1e-12, determinism, linearity."""
    from cannoles_b200.workloads import first_system, make_config
    nls, method, _ = make_config("c3", None)
    s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                                                         ordering=3, refine_steps=3, shift_retries=True))
    B, N = s.LDLT, s.LDLT.N
    assert N == 7_450_007
    ok, rho, tries = _rho_retry(B, s, nls)
    assert ok and B.last_inertia == (nls.nvar, 0, nls.nequ + nls.ncon, False)
    assert tries >= 1 and B.n_shift == tries            # retries went through the device-side shift
    d1 = B.factor.d
    assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    assert np.array_equal(d1, B.factor.d)
    x1, x2, x3 = np.zeros(N), np.zeros(N), np.zeros(N)
    b2 = np.random.default_rng(9).standard_normal(N)
    B.solve_ldl(rhs, x1)
    assert B.last_relres <= ec.RESID_TOL
    B.solve_ldl(b2, x2)
    assert B.last_relres <= ec.RESID_TOL
    B.solve_ldl(2.5 * rhs + b2, x3)
    assert np.linalg.norm(x3 - (2.5 * x1 + x2)) <= 1e-7 * np.linalg.norm(x3)
    B.close()


@pytest.mark.parametrize("ordering", [0, 3])
def test_c3_slice_against_oracle_with_rho_retry(ctor, oracle_cls, ordering):
    """A 5000-camera slice of config 3 (N = 745 007, nnz(L) ~ 14 M): the rho = 0 breakdown, the rho0
    retry, then the oracle on the same order: nnz(L), CSC values, inertia equal; pivots within 1e-7
    (rho0 = 6e-6 on a zero (1,1) block: pivots span 1e-6 .. 1e+3), step within 1e-7."""
    from cannoles_b200.workloads import first_system, make_config
    nls, method, _ = make_config("c3", 5000)
    s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                                                         ordering=ordering, refine_steps=3))
    B = s.LDLT
    ok0 = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    O = oracle_cls(B.N, s.rows, s.cols, s.vals, perm=B.perm)
    assert O.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS) == ok0
    assert B.stats()["nnzL"] == O.nnzL
    ok, rho, tries = _rho_retry(B, s, nls) if not ok0 else (True, 0.0, 0)
    assert ok and B.last_inertia == (nls.nvar, 0, nls.nequ + nls.ncon, False)
    assert O.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    assert np.array_equal(B.nzval, O.nzval)
    assert B.last_inertia[:3] == O.inertia(EPS)
    dB, dO = B.factor.d, O.factor.d
    assert np.max(np.abs(dB - dO) / np.abs(dO)) < 1e-7
    d, do = np.zeros(B.N), np.zeros(B.N)
    B.solve_ldl(rhs, d)
    O.solve_ldl(rhs, do)
    assert B.last_relres <= ec.RESID_TOL
    assert np.linalg.norm(d - do) <= 1e-7 * np.linalg.norm(do)
    B.close()


def test_small_delta_needs_refinement_and_reaches_the_bar(ctor, oracle_cls):
    """delta = sqrt(eps) (the smallest value the loop ever uses, reference/src/CaNNOLeS.jl:52, 615):
    pivots span 1e-8 .. 1e+1, the pivot-free factorization loses digits and the adaptive
    refinement has to recover them: relative residual <= 1e-12 (north_star)."""
    nv, ne, nc = 300, 400, 120
    N, r, c, v = random_kkt(nv, ne, nc, 0.03, 77, delta=1.4901161193847656e-08)
    B = ctor(N, r, c, v, nvar=nv, nequ=ne, ncon=nc, refine_steps=3)
    assert B.try_to_factorize(v, nv, ne, nc, EPS)
    O = oracle_cls(N, r, c, v, perm=B.perm)
    assert O.try_to_factorize(v, nv, ne, nc, EPS)
    assert B.last_inertia[:3] == O.inertia(EPS) == (nv, 0, ne + nc)
    rhs = np.random.default_rng(8).standard_normal(N)
    d = np.zeros(N)
    B.solve_ldl(rhs, d)
    assert B.last_relres <= ec.RESID_TOL
    assert np.linalg.norm(O.matvec(d) + rhs) <= ec.RESID_TOL * np.linalg.norm(rhs)
    assert 1 <= B.last_sweeps <= 4


def test_gauss_newton_breakdown_takes_the_rho_retry_like_the_reference(ctor, oracle_cls):
    """C3 slice (Gauss-Newton: zero (1,1) block).  With a fill-reducing order the rho = 0 matrix can
    hit zero pivots; the caller protocol of newton_system! (rho0, then x100) must end in the expected
    inertia and the retried factorization must agree with the oracle on the same order."""
    from cannoles_b200.workloads import first_system, make_config
    nls, method, _ = make_config("c3", 300)
    s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon, ordering=3))
    B = s.LDLT
    ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    rho, tries = 0.0, 0
    while not ok and tries < 6:
        rho = EPS ** (1.0 / 3.0) if rho == 0.0 else 100.0 * rho
        s.vals[len(s.vals) - nls.nvar:] = rho
        ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
        tries += 1
    assert ok and B.last_inertia == (nls.nvar, 0, nls.nequ + nls.ncon, False)
    O = oracle_cls(B.N, s.rows, s.cols, s.vals, perm=B.perm)
    assert O.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    assert B.last_inertia[:3] == O.inertia(EPS)
    d, do = np.zeros(B.N), np.zeros(B.N)
    B.solve_ldl(rhs, d)
    O.solve_ldl(rhs, do)
    assert B.last_relres <= ec.RESID_TOL
    assert np.linalg.norm(d - do) <= 1e-7 * np.linalg.norm(do)


def test_inspection_entry_points(ctor, gpu_lib):
    """b2_front_sizes / b2_profile / b2_host_register: shapes and basic invariants."""
    import ctypes as C
    from cannoles_b200 import _capi
    N, r, c, v = random_kkt(120, 160, 40, 0.05, 78)
    B = ctor(N, r, c, v, nvar=120, nequ=160, ncon=40)
    assert B.try_to_factorize(v, 120, 160, 40, EPS)
    d = np.zeros(N)
    B.solve_ldl(np.ones(N), d)
    ns = B.stats()["nsuper"]
    w = np.zeros(ns, dtype=np.int32); m = np.zeros(ns, dtype=np.int32); lv = np.zeros(ns, dtype=np.int32)
    assert gpu_lib.b2_front_sizes(B._h, ns, w.ctypes.data_as(_capi.p32), m.ctypes.data_as(_capi.p32),
                                  lv.ctypes.data_as(_capi.p32)) == 0
    assert w.sum() == N and (m >= w).all() and lv.max() + 1 == B.stats()["nlevels"]
    MAXN = 4096
    kinds = np.zeros(MAXN, dtype=np.int32); cls = np.zeros(MAXN, dtype=np.int32)
    counts = np.zeros(MAXN, dtype=np.int32); ms = np.zeros(MAXN); n = C.c_int()
    pi = C.POINTER(C.c_int)
    for which in (0, 1):
        assert gpu_lib.b2_profile(B._h, which, MAXN, kinds.ctypes.data_as(pi), cls.ctypes.data_as(pi),
                                  counts.ctypes.data_as(pi), ms.ctypes.data_as(_capi.pd), C.byref(n)) == 0
        assert n.value > 0 and (ms[:n.value] >= 0).all()
    d2 = np.zeros(N)
    B.solve_ldl(np.ones(N), d2)                    # the profile replays must not disturb the factor
    assert np.array_equal(d, d2)
    buf = np.zeros(1 << 16)
    assert gpu_lib.b2_host_register(buf.ctypes.data_as(C.c_void_p), buf.nbytes) == 0
    assert gpu_lib.b2_host_unregister(buf.ctypes.data_as(C.c_void_p)) == 0


def test_factorization_is_bit_reproducible_over_many_runs(ctor, monkeypatch):
    """300 factorizations of the 5000-camera slice of config 3 with EVERY tiled front on the dataflow
    kernel (B2_DAG_MIN_NP = 1): pivots and inertia identical bit for bit.  Round 2 found a write-after-
    read race on the child descriptors of k_front_dag this way (one run in ~100 differed)."""
    from cannoles_b200.workloads import first_system, make_config
    monkeypatch.setenv("B2_DAG_MIN_NP", "1")
    nls, method, _ = make_config("c3", 5000)
    s, rhs = first_system(nls, method, functools.partial(ctor, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                                                         shift_retries=False))
    B = s.LDLT
    ok0 = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    d0, i0 = B.factor.d.copy(), B.last_inertia
    for _ in range(300):
        assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS) == ok0
        assert B.last_inertia == i0
        assert np.array_equal(B.factor.d, d0, equal_nan=True)
    B.close()
