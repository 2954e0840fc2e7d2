"""`cannoles` wall time with the B200 backend (and, where it finishes quickly, with the CPU oracle
on the same elimination order): the "cannoles wall time" half of BASELINE.json's metric.
The optimisation loop is the restated reference loop (cannoles_b200/solver.py, host Python/numpy:
model callbacks, CGLS, line search); the backend time is what this repository accelerates.
usage: loop_bench.py [c2|c4|c1] [size] [--oracle]   -> one JSON line per run"""
import functools
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cannoles_b200 import CaNNOLeSSolver, solve  # noqa: E402
from cannoles_b200.linsolve import B200Struct  # noqa: E402
from cannoles_b200.workloads import make_config  # noqa: E402


class Timed:
    """Wraps a backend and accumulates the wall time spent inside its verbs."""

    def __init__(self, inner):
        self.inner, self.t_fact, self.t_solve, self.nf, self.ns = inner, 0.0, 0.0, 0, 0
        self.factor = inner.factor

    def get_vals(self):
        return self.inner.get_vals()

    def try_to_factorize(self, *a):
        t = time.perf_counter()
        ok = self.inner.try_to_factorize(*a)
        self.t_fact += time.perf_counter() - t
        self.nf += 1
        return ok

    def solve_ldl(self, rhs, d):
        t = time.perf_counter()
        ok = self.inner.solve_ldl(rhs, d)
        self.t_solve += time.perf_counter() - t
        self.ns += 1
        return ok


def run(cfg, size, use_oracle, perm=None):
    nls, method, desc = make_config(cfg, size)
    holder = {}

    def ctor(N, rows, cols, vals):
        if use_oracle:
            from oracle import LDLFactStruct
            inner = LDLFactStruct(N, rows, cols, vals, perm=perm)
        else:
            inner = B200Struct(N, rows, cols, vals, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                               ordering=3 if cfg == "c3" else 0)
        holder["b"] = Timed(inner)
        return holder["b"]

    t0 = time.perf_counter()
    s = CaNNOLeSSolver(nls, linsolve=ctor, method=method)
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = solve(s, nls, max_time=3600.0)
    wall = time.perf_counter() - t0
    b = holder["b"]
    rec = {"config": f"{cfg}: {desc}", "backend": "oracle (CPU port)" if use_oracle else "b200",
           "N": nls.nvar + nls.nequ + nls.ncon, "status": st.status, "iter": st.iter,
           "nfact": st.solver_specific["nfact"], "nlinsolve": st.solver_specific["nlinsolve"],
           "objective": st.objective, "primal_feas": st.primal_feas, "dual_feas": st.dual_feas,
           "wall_s": wall, "setup_s": t_setup, "backend_factor_s": b.t_fact, "backend_solve_s": b.t_solve,
           "host_loop_s": wall - b.t_fact - b.t_solve}
    print(json.dumps(rec), flush=True)
    return st, (None if use_oracle else b.inner.perm)


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    size = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else None
    st_b, perm = run(cfg, size, False)
    if "--oracle" in sys.argv:
        st_o, _ = run(cfg, size, True, perm=perm)
        same = (st_b.iter == st_o.iter and st_b.solver_specific["nfact"] == st_o.solver_specific["nfact"]
                and np.linalg.norm(st_b.solution - st_o.solution) <= 1e-8 * max(1.0, np.linalg.norm(st_o.solution)))
        print(json.dumps({"same_iterations_nfact_and_solution_1e-8": bool(same)}), flush=True)
