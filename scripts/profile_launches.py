"""Warm-cache per-launch device times of one factorization and one solve sweep (b2_profile).
usage: profile_launches.py [c4|c2] [size] [ordering] -> table on stdout"""
import collections
import ctypes as C
import functools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cannoles_b200 import _capi  # noqa: E402
from cannoles_b200.linsolve import B200Struct  # noqa: E402
from cannoles_b200.workloads import first_system, make_config  # noqa: E402

KIND = ["front_small", "assemble_large", "diag_writeback", "trsm", "update", "fwd", "bwd", "fwd_big", "bwd_big", "front_tiny", "fwd_tiny", "bwd_tiny", "diag", "dag", "linv"]
EPS = 2.0 ** -52
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
size = int(sys.argv[2]) if len(sys.argv) > 2 else None
ordering = int(sys.argv[3]) if len(sys.argv) > 3 else 0
verbose = len(sys.argv) > 4
nls, method, desc = make_config(cfg, size)
ctor = functools.partial(B200Struct, ordering=ordering, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon,
                         refine_steps=0, shift_retries=False)
s, rhs = first_system(nls, method, ctor)
B = s.LDLT
d = np.zeros(B.N)
if not B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS):
    s.vals[len(s.vals) - nls.nvar:] = EPS ** (1.0 / 3.0)      # the reference's first rho retry
for _ in range(2):
    assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    B.solve_ldl(rhs, d)
print("graph timings (ms):", B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS) and B.timings())
B.solve_ldl(rhs, d)
print("graph timings (ms):", B.timings())
lib = _capi.load()
MAXN = 100000
for which, name in ((0, "factorization"), (1, "solve sweep")):
    kinds = np.zeros(MAXN, dtype=np.int32); cls = np.zeros(MAXN, dtype=np.int32)
    counts = np.zeros(MAXN, dtype=np.int32); ms = np.zeros(MAXN); n = C.c_int()
    pi = C.POINTER(C.c_int)
    rc = lib.b2_profile(B._h, which, MAXN, kinds.ctypes.data_as(pi), cls.ctypes.data_as(pi),
                        counts.ctypes.data_as(pi), ms.ctypes.data_as(_capi.pd), C.byref(n))
    assert rc == 0, _capi.last_error(lib)
    n = n.value
    agg = collections.defaultdict(lambda: [0, 0, 0.0])
    for i in range(n):
        key = (KIND[kinds[i]], int(cls[i]))
        agg[key][0] += 1; agg[key][1] += int(counts[i]); agg[key][2] += ms[i]
    tot = ms[:n].sum()
    print(f"# {name}: {n} launches, {tot:.3f} ms (events between launches, warm cache, no graph)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"  {k[0]:16s} cls/mode {k[1]}  launches {v[0]:5d}  CTAs {v[1]:9d}  {v[2]:8.3f} ms  {100 * v[2] / tot:5.1f}%")
    if verbose:
        for i in range(n):
            print(f"    {i:4d} {KIND[kinds[i]]:16s} {int(cls[i])} ctas={int(counts[i]):7d} {ms[i] * 1e3:9.1f} us")
