"""Parity checks of the batched engine (one CTA per KKT system) shared by the CPU-emulator
tests and the GPU tests: every instance against the oracle on the same elimination order."""
import numpy as np

from cannoles_b200.batched import B200BatchStruct
from cannoles_b200.solver import ParamCaNNOLeS, newton_system
from tests.problems import EPS, random_kkt

RESID_TOL = 1e-12


def random_batch(nv, ne, nc, dens, seed, batch, indefinite=()):
    """`batch` value sets on ONE random KKT pattern; instances in `indefinite` get a (1,1) block
    with negative curvature so that the rho = 0 factorization has the wrong inertia."""
    N, rows, cols, v0 = random_kkt(nv, ne, nc, dens, seed)
    rng = np.random.default_rng(seed + 100)
    vals = np.empty((batch, len(v0)))
    nH = int(np.argmax(rows > nv))           # entries of the H segment come first
    for b in range(batch):
        v = v0 * (1.0 + 0.3 * rng.standard_normal(len(v0)))
        v[-(nv + nc + ne):-(nv + nc)] = -1.0          # -I segment
        v[-(nv + nc):-nv] = -0.1 * (1 + b % 3)        # -delta segment
        v[-nv:] = 0.0                                 # rho segment
        if b in indefinite:
            diag = (rows[:nH] == cols[:nH])
            v[:nH][diag] = -np.abs(v[:nH][diag]) * 2.0
        vals[b] = v
    rhs = rng.standard_normal((batch, N))
    return N, rows, cols, vals, rhs


def check_batch_against_oracle(lib, oracle_cls, nv, ne, nc, dens, seed, batch, ordering=3, tol=1e-9):
    N, rows, cols, vals, rhs = random_batch(nv, ne, nc, dens, seed, batch)
    Bt = B200BatchStruct(N, rows, cols, batch, nv, ne, nc, ordering=ordering, _lib=lib)
    ok = Bt.try_to_factorize(vals, EPS)
    d = np.zeros((batch, N))
    Bt.solve_ldl(rhs, d)
    d2 = np.zeros((batch, N))
    Bt2 = B200BatchStruct(N, rows, cols, batch, nv, ne, nc, ordering=ordering, _lib=lib)
    ok2 = Bt2.factor_solve(vals, rhs, d2)                 # fused verb == the two separate verbs
    assert np.array_equal(ok, ok2) and np.array_equal(d, d2)
    O = oracle_cls(N, rows, cols, vals[0].copy(), perm=Bt.perm)
    for b in range(batch):
        ok_o = O.try_to_factorize(vals[b], nv, ne, nc, EPS)
        assert bool(ok[b]) == ok_o
        assert (Bt.npos[b], Bt.nzero[b], Bt.nneg[b]) == O.inertia(EPS)   # bit-exact counts
        dB, dO = Bt.d_of(b), O.factor.d
        assert np.max(np.abs(dB - dO) / np.abs(dO)) < tol
        xo = np.zeros(N)
        O.solve_ldl(rhs[b], xo)
        assert np.linalg.norm(d[b] - xo) <= tol * np.linalg.norm(xo)
        assert np.linalg.norm(O.matvec(d[b]) + rhs[b]) <= RESID_TOL * np.linalg.norm(rhs[b])
    Bt.close(); Bt2.close()


def check_batched_newton_system(lib, oracle_cls, nv=24, ne=30, nc=6, batch=7, seed=41):
    """The masked rho driver: every instance must see the rho sequence, nfact and step that the
    restated newton_system! gives it alone (oracle backend, same order)."""
    bad = (1, 4)
    N, rows, cols, vals, rhs = random_batch(nv, ne, nc, 0.15, seed, batch, indefinite=bad)
    Bt = B200BatchStruct(N, rows, cols, batch, nv, ne, nc, _lib=lib)
    params = ParamCaNNOLeS()
    rho_old = np.zeros(batch)
    rho_old[4] = 1e-3            # one instance with a previous successful rho (kappa_dec branch)
    d = np.zeros((batch, N))
    vb = vals.copy()
    active = np.ones(batch, dtype=bool)
    active[6] = False            # a converged instance: must be left alone
    d[6] = 7.0
    succ, rho, rho_new, nfact = Bt.newton_system(d, rhs, vb, rho_old, params, active)
    assert nfact[6] == 0 and np.all(d[6] == 7.0) and not succ[6]
    for b in range(batch):
        if not active[b]:
            continue
        vo = vals[b].copy()
        O = oracle_cls(N, rows, cols, vo, perm=Bt.perm)
        do = np.zeros(N)
        _, s_o, rho_o, rho_old_o, nfact_o = newton_system(do, nv, ne, nc, rhs[b], vo, O, rho_old[b], params)
        assert bool(succ[b]) == bool(s_o)
        assert rho[b] == rho_o and rho_new[b] == rho_old_o and nfact[b] == nfact_o, (b, rho[b], rho_o, nfact[b], nfact_o)
        assert np.array_equal(vb[b], vo)                       # rho segment left as the reference leaves it
        assert (b in bad) == (nfact_o > 1)
        assert np.linalg.norm(d[b] - do) <= 1e-8 * np.linalg.norm(do)
    Bt.close()
