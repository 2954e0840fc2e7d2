"""Run one warm-up + one profiled factorize/solve of a single-system config (for ncu).
usage: profile_single.py <c2|c4> [size] [ordering]"""
import functools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cannoles_b200.linsolve import B200Struct  # noqa: E402
from cannoles_b200.models import ExtRosenbrockLinEq, PoissonParamEst  # noqa: E402
from scripts.gpu_check import first_system  # noqa: E402

EPS = 2.0 ** -52
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
size = int(sys.argv[2]) if len(sys.argv) > 2 else (512 if cfg == "c4" else 100_000)
ordering = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
nls, method = (PoissonParamEst(size), "Newton") if cfg == "c4" else (ExtRosenbrockLinEq(size), "Newton_noFHess")
ctor = functools.partial(B200Struct, ordering=ordering, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                         refine_steps=0, shift_retries=False)
s, rhs = first_system(nls, method, ctor)
B = s.LDLT
d = np.zeros(B.N)
for it in range(reps):
    ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    tf = B.timings()
    B.solve_ldl(rhs, d)
    ts = B.timings()
    print(it, ok, B.last_inertia, "factor ms %.3f solve ms %.3f" % (tf["factor"], ts["solve"]), flush=True)
print(B.stats())
