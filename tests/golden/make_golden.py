"""Generates tests/golden/*.npz.

The reference (Julia) cannot run in this image and its tests hold no vectors at the
factorization boundary, so these fixtures are produced by the CPU oracle (oracle/ldl_oracle.c)
and cross-checked against a dense LAPACK solve before being written; the hand-derived MGH01CON
vectors of SURVEY App. D are stored verbatim.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import LDLFactStruct  # noqa: E402
from tests.problems import EPS, dense_from_coo, random_kkt  # noqa: E402


def write_ref_input(name, N, rows, cols, vals, nvar, nequ, ncon, rhs):
    """The same inputs as plain text for oracle/gen_ref_vectors.jl (Julia stdlib only): line 1
    `N nnz nvar nequ ncon`, nnz lines `row col val` (1-based), N lines rhs."""
    os.makedirs(os.path.join(HERE, "ref_inputs"), exist_ok=True)
    with open(os.path.join(HERE, "ref_inputs", name + ".txt"), "w") as f:
        f.write(f"{N} {len(vals)} {nvar} {nequ} {ncon}\n")
        for r, c, v in zip(rows, cols, vals):
            f.write(f"{int(r)} {int(c)} {float(v):.17g}\n")
        for v in rhs:
            f.write(f"{float(v):.17g}\n")


def one(name, N, rows, cols, vals, nvar, nequ, ncon, perm=None):
    L = LDLFactStruct(N, rows, cols, vals, perm=perm)
    ok = L.try_to_factorize(vals, nvar, nequ, ncon, EPS)
    rhs = np.random.default_rng(11).standard_normal(N)
    d = np.zeros(N)
    if ok:
        L.solve_ldl(rhs, d)
        K = dense_from_coo(N, rows, cols, vals)
        assert np.allclose(d, -np.linalg.solve(K, rhs), rtol=1e-9, atol=1e-11)
        ev = np.linalg.eigvalsh(K)
        assert L.inertia(EPS) == (int((ev > 0).sum()), 0, int((ev < 0).sum()))
    write_ref_input(name, N, rows, cols, vals, nvar, nequ, ncon, rhs)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), N=N, rows=rows, cols=cols, vals=vals,
                        dims=np.array([nvar, nequ, ncon]), perm=L.perm, colptr=L.colptr,
                        rowval=L.rowval, nzval=L.nzval, D=L.factor.d, inertia=np.array(L.inertia(EPS)),
                        ok=ok, rhs=rhs, d=d)


if __name__ == "__main__":
    rows = np.array([1, 1, 2, 2, 3, 4, 4, 5, 3, 4, 5, 1, 2], dtype=np.int64)
    cols = np.array([1, 1, 1, 2, 1, 1, 2, 1, 3, 4, 5, 1, 2], dtype=np.int64)
    vals = np.array([88., 0, 0, 0, -1, 24, 10, 1, -1, -1, -0.1, 0, 0])
    one("mgh01con_first_kkt", 5, rows, cols, vals, 2, 2, 1, perm=[4, 2, 0, 3, 1])
    for i, (nv, ne, nc, dens) in enumerate([(20, 30, 5, 0.2), (60, 80, 20, 0.08), (150, 200, 40, 0.03)]):
        N, r, c, v = random_kkt(nv, ne, nc, dens, 100 + i)
        one(f"random_kkt_{i}", N, r, c, v, nv, ne, nc)
