"""Repeat the factorization of one KKT system and compare the pivots bit for bit with the first run.
usage: determinism_loop.py [c4|c3] [size] [reps]"""
import functools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from cannoles_b200.linsolve import B200Struct
from cannoles_b200.workloads import first_system, make_config
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
size = int(sys.argv[2]) if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
EPS = 2.0 ** -52
nls, method, _ = make_config(cfg, size)
s, rhs = first_system(nls, method, functools.partial(B200Struct, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon, shift_retries=False))
B = s.LDLT
ok0 = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
d0 = B.factor.d.copy()
i0 = B.last_inertia
import ctypes as C
from cannoles_b200 import _capi
lib = _capi.load()
st = B.stats()
ns = int(st["nsuper"])
wd = np.zeros(ns, dtype=np.int32); od = np.zeros(ns, dtype=np.int32); lv = np.zeros(ns, dtype=np.int32)
p32 = C.POINTER(C.c_int32)
lib.b2_front_sizes(B._h, ns, wd.ctypes.data_as(p32), od.ctypes.data_as(p32), lv.ctypes.data_as(p32))
starts = np.concatenate([[0], np.cumsum(wd)])
print("nlevels", st["nlevels"], "nsuper", ns, "max_front", st["max_front"])
bad = 0
for r in range(reps):
    ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    d = B.factor.d
    if ok != ok0 or B.last_inertia != i0 or not np.array_equal(d, d0, equal_nan=True):
        bad += 1
        idx = np.flatnonzero(~((d == d0) | (np.isnan(d) & np.isnan(d0))))
        fs = sorted(set(int(np.searchsorted(starts, i, side="right") - 1) for i in idx[:2000]))
        print("   fronts:", [(f, int(starts[f]), int(wd[f]), int(od[f]), int(lv[f])) for f in fs[:8]], "(front, first column, width, order, level)")
        print("rep", r, "MISMATCH ok", ok, "inertia", B.last_inertia, "ndiff", idx.size, "first", idx[:5], d[idx[:3]], d0[idx[:3]], flush=True)
print(cfg, size, "env", {k: v for k, v in os.environ.items() if k.startswith("B2_")}, "ok0", ok0, i0, "mismatches", bad, "of", reps)
