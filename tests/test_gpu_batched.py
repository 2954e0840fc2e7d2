"""-m gpu: the batched engine on the B200 through the C ABI, against the oracle."""
import numpy as np
import pytest

from tests import batched_checks as bc
from tests.problems import EPS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ordering", [0, 3])
def test_batched_against_oracle(gpu_lib, oracle_cls, ordering):
    bc.check_batch_against_oracle(gpu_lib, oracle_cls, 60, 90, 20, 0.08, 61, batch=40, ordering=ordering)


def test_batched_edge_shapes(gpu_lib, oracle_cls):
    bc.check_batch_against_oracle(gpu_lib, oracle_cls, 9, 14, 0, 0.3, 52, batch=5)
    bc.check_batch_against_oracle(gpu_lib, oracle_cls, 2, 2, 1, 1.0, 53, batch=3, ordering=1)
    bc.check_batch_against_oracle(gpu_lib, oracle_cls, 70, 110, 28, 0.9, 54, batch=4)   # N = 208, dense


def test_batched_newton_system_matches_per_instance_driver(gpu_lib, oracle_cls):
    bc.check_batched_newton_system(gpu_lib, oracle_cls, nv=40, ne=60, nc=12, batch=33, seed=62)


def test_config5_instances_against_oracle(gpu_lib, oracle_cls):
    """BASELINE config 5 shape (n=64, m=128, 16 constraints): 96 instances, each against the oracle."""
    from cannoles_b200.batched import B200BatchStruct
    from cannoles_b200.workloads import dense_batch_systems
    nb = 96
    s, vals, rhs = dense_batch_systems(range(nb))
    nv, ne, nc = s.nvar, s.nequ, s.ncon
    N = nv + ne + nc
    Bt = B200BatchStruct(N, s.rows, s.cols, nb, nv, ne, nc)
    d = np.zeros((nb, N))
    ok = Bt.factor_solve(vals, rhs, d)
    assert ok.all()
    O = oracle_cls(N, s.rows, s.cols, vals[0].copy(), perm=Bt.perm)
    for b in range(nb):
        assert O.try_to_factorize(vals[b], nv, ne, nc, EPS)
        assert (Bt.npos[b], Bt.nzero[b], Bt.nneg[b]) == O.inertia(EPS) == (nv, 0, ne + nc)
        xo = np.zeros(N)
        O.solve_ldl(rhs[b], xo)
        assert np.linalg.norm(d[b] - xo) <= 1e-9 * np.linalg.norm(xo)
        assert np.linalg.norm(O.matvec(d[b]) + rhs[b]) <= 1e-12 * np.linalg.norm(rhs[b])
    # determinism: a second launch gives bit-identical steps
    d2 = np.zeros((nb, N))
    Bt.factor_solve(vals, rhs, d2)
    assert np.array_equal(d, d2)
    Bt.close()


def test_config5_full_batch_inertia_against_oracle(gpu_lib, oracle_cls):
    """BASELINE config 5 at its full size: all 8192 instances, the inertia triple of each equal to the
    oracle's on the same elimination order (reference/src/solver_types.jl:90-96), and the step of every
    16th instance within 1e-9."""
    from cannoles_b200.batched import B200BatchStruct
    from cannoles_b200.workloads import dense_batch_systems
    nb = 8192
    s, vals, rhs = dense_batch_systems(range(nb))
    nv, ne, nc = s.nvar, s.nequ, s.ncon
    N = nv + ne + nc
    Bt = B200BatchStruct(N, s.rows, s.cols, nb, nv, ne, nc)
    d = np.zeros((nb, N))
    ok = Bt.factor_solve(vals, rhs, d)
    O = oracle_cls(N, s.rows, s.cols, vals[0].copy(), perm=Bt.perm)
    for b in range(nb):
        ok_o = O.try_to_factorize(vals[b], nv, ne, nc, EPS)
        assert bool(ok[b]) == ok_o
        assert (Bt.npos[b], Bt.nzero[b], Bt.nneg[b]) == O.inertia(EPS)
        if ok_o and b % 16 == 0:
            xo = np.zeros(N)
            O.solve_ldl(rhs[b], xo)
            assert np.linalg.norm(d[b] - xo) <= 1e-9 * np.linalg.norm(xo)
    Bt.close()


def test_full_batch_properties(gpu_lib):
    """A large batch (2048 instances): expected inertia everywhere, residual through linearity
    (solve(2 rhs) == 2 solve(rhs) on the stored factors), masks respected."""
    from cannoles_b200.batched import B200BatchStruct
    from cannoles_b200.workloads import dense_batch_systems
    nb = 2048
    s, vals, rhs = dense_batch_systems(range(64))
    reps = nb // 64
    vals = np.ascontiguousarray(np.tile(vals, (reps, 1)) * (1 + 1e-3 * np.arange(nb)[:, None] / nb))
    rhs = np.ascontiguousarray(np.tile(rhs, (reps, 1)))
    nv, ne, nc = s.nvar, s.nequ, s.ncon
    N = nv + ne + nc
    Bt = B200BatchStruct(N, s.rows, s.cols, nb, nv, ne, nc)
    ok = Bt.try_to_factorize(vals, EPS)
    assert ok.all() and (Bt.npos == nv).all() and (Bt.nneg == ne + nc).all()
    d1, d2 = np.zeros((nb, N)), np.full((nb, N), 3.0)
    Bt.solve_ldl(rhs, d1)
    act = np.zeros(nb, dtype=np.uint8)
    act[::2] = 1
    Bt.solve_ldl(2.0 * rhs, d2, active=act)
    assert np.allclose(d2[::2], 2.0 * d1[::2], rtol=1e-12, atol=0)
    assert np.all(d2[1::2] == 3.0)
    Bt.close()
