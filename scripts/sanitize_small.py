"""Small cases for compute-sanitizer (memcheck / racecheck): every kernel family once.
usage: compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import functools
import os
import sys

import numpy as np

os.environ.setdefault("B2_SMALL_MAX_M", "16")    # fronts of order > 16 take the tiled path
os.environ.setdefault("B2_SOLVE_BIG_M", "24")    # fronts of order > 24 take the multi-CTA solves
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cannoles_b200.batched import B200BatchStruct  # noqa: E402
from cannoles_b200.linsolve import B200Struct  # noqa: E402
from cannoles_b200.workloads import dense_batch_systems, first_system, make_config  # noqa: E402

EPS = 2.0 ** -52
# the last case has fronts with three pivot blocks and runs every tiled front through the dataflow
# kernel (k_front_dag: plain, chain and ypre tasks); the first two take the k_trsm / k_update chain
for cfg, size, dag_min_np in (("c4", 14, "0"), ("c2", 300, "0"), ("c4", 60, "1000000000"), ("c4", 60, "40")):
    os.environ["B2_DAG_LEVEL_MAX"] = dag_min_np   # 0: launch chain only; huge: every front through k_front_dag; 40: the top levels
    nls, method, _ = make_config(cfg, size)
    ctor = functools.partial(B200Struct, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon, refine_steps=1,
                             refine_tol=0.0, shift_retries=True)
    s, rhs = first_system(nls, method, ctor)
    B = s.LDLT
    ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    s.vals[-nls.nvar:] = 1e-6
    ok2 = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)     # device-side shift path
    d = np.zeros(B.N)
    B.solve_ldl(rhs, d)
    st = B.stats()
    print(cfg, size, "N", B.N, "ok", ok, ok2, "relres", B.last_relres, "n_large", st["n_large"], "max_front", st["max_front"], "max_width", st["max_width"], "dag_level_max", dag_min_np)
    B.close()
nb = 3
s, vals, rhs = dense_batch_systems(range(nb))
Bt = B200BatchStruct(208, s.rows, s.cols, nb, 64, 128, 16)
d = np.zeros((nb, 208))
ok = Bt.factor_solve(vals, rhs, d)
ok3 = Bt.try_to_factorize(vals)
Bt.solve_ldl(rhs, d)
print("batched", ok, ok3, float(np.abs(d).max()))
Bt.close()

# the device-resident batched loop (k_nls_dense: TMA staging, CGLS, line search via a far start)
from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models  # noqa: E402
mod = pack_dense_models(range(2))
mod["x0"][1] *= 10.0
S = B200BatchNLS(2)
rec = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
print("nls status", rec[:, 0], "iter", rec[:, 1], "nfact", rec[:, 2], "nbk", rec[:, 4])
S.close()
