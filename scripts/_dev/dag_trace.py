"""Developer aid: task trace of k_front_dag on C4 (needs scripts/_dev/libb2_timing.so, built by
build_timing.sh).  Prints, for the biggest fronts, when every pivot block was factored and how the
time of its critical path splits (diag factorization / substitution / last update step / waits)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200 import _capi
lib = _capi.bind_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb2_timing.so"))
_capi._LIB = lib
from cannoles_b200.linsolve import B200Struct
from cannoles_b200.models import PoissonParamEst
from cannoles_b200.workloads import first_system
import functools
EPS = 2.0 ** -52
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nls = PoissonParamEst(size)
ctor = functools.partial(B200Struct, ordering=0, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon, refine_steps=1)
s, rhs = first_system(nls, "Newton", ctor)
B = s.LDLT
for _ in range(3):
    assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
print(B.timings())
NT, W = 1 << 16, 12
out = (C.c_longlong * (W * NT))()
lib.b2_debug_dag_trace.argtypes = [C.POINTER(C.c_longlong), C.c_longlong]
assert lib.b2_debug_dag_trace(out, W * NT) == 0
tr = np.frombuffer(out, dtype=np.int64).reshape(NT, W)
tr = tr[tr[:, 2] > 0]
print("tasks traced", len(tr))
# slots: 0 id, 1 sm, 2 start, 3 updates done, 4 flag of the pivot block seen, 5 L loaded, 6 substitution done,
#        7 LDL^T start, 8 LDL^T end, 9 end
front = tr[:, 0] >> 32; chain = (tr[:, 0] >> 30) & 1; I = (tr[:, 0] >> 15) & 0x7fff; J = tr[:, 0] & 0x7fff
t0 = tr[:, 2].min()
us = lambda a, b: (a - b) / 1e3
ids, cnt = np.unique(front, return_counts=True)
for f in ids[np.argsort(-cnt)[:3]]:
    sel = front == f
    a = tr[sel]; ii = I[sel]; jj = J[sel]; ch = chain[sel]
    print(f"front {f}: {sel.sum()} tasks, span {us(a[:, 9].max(), a[:, 2].min()):.1f} us, starts at {us(a[:, 2].min(), t0):.1f} us")
    prev = None
    for k in np.argsort(jj):
        if ii[k] != jj[k] or a[k, 8] == 0: continue
        d = a[k]
        line = f"  J={jj[k]:3d} sm={d[1]:3d} start={us(d[2], t0):8.1f} upd_done={us(d[3], t0):8.1f}"
        if ch[k]:
            line += f" flag_seen={us(d[4], t0):8.1f} loadL={us(d[5], d[4]):4.1f} subst={us(d[6], d[5]):4.1f} upd+store={us(d[7], d[6]):4.1f}"
        line += f" ldlt={us(d[8], d[7]):5.1f} post={us(d[9], d[8]):4.1f} end={us(d[9], t0):8.1f}"
        if prev is not None: line += f"  since prev {us(d[9], prev):5.1f}"
        print(line)
        prev = d[9]
