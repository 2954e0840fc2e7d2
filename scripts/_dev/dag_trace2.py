"""Developer aid: per-level timeline of the cross-level k_front_dag launch on C4 (needs
scripts/_dev/libb2_timing.so, built by build_timing.sh)."""
import ctypes as C, os, sys, functools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200 import _capi
lib = _capi.bind_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb2_timing.so"))
_capi._LIB = lib
from cannoles_b200.linsolve import B200Struct
from cannoles_b200.models import PoissonParamEst
from cannoles_b200.workloads import first_system
EPS = 2.0 ** -52
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nls = PoissonParamEst(size)
ctor = functools.partial(B200Struct, ordering=0, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon, refine_steps=1)
s, rhs = first_system(nls, "Newton", ctor)
B = s.LDLT
for _ in range(3):
    assert B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
print(B.timings())
ns = B.stats()["nsuper"]
w = np.zeros(ns, np.int32); m = np.zeros(ns, np.int32); lv = np.zeros(ns, np.int32)
lib.b2_front_sizes(B._h, ns, w.ctypes.data_as(_capi.p32), m.ctypes.data_as(_capi.p32), lv.ctypes.data_as(_capi.p32))
NT, W = 1 << 16, 12
out = (C.c_longlong * (W * NT))()
lib.b2_debug_dag_trace.argtypes = [C.POINTER(C.c_longlong), C.c_longlong]
assert lib.b2_debug_dag_trace(out, W * NT) == 0
tr = np.frombuffer(out, dtype=np.int64).reshape(NT, W)
tr = tr[tr[:, 2] > 0]
print("tasks traced", len(tr))
front = tr[:, 0] >> 32; chain = (tr[:, 0] >> 30) & 1; I = (tr[:, 0] >> 15) & 0x7fff; J = tr[:, 0] & 0x7fff
t0 = tr[:, 2].min(); tend = tr[:, 9].max()
us = lambda a: (a - t0) / 1e3
print(f"launch span {us(tend):.1f} us; CTA-time {((tr[:, 9] - tr[:, 2]).sum()) / 1e3:.0f} us = {((tr[:, 9] - tr[:, 2]).sum()) / (tend - t0) :.1f} CTAs busy on average")
L = lv[front]
for l in np.unique(L):
    q = L == l
    a = tr[q]
    dur = (a[:, 9] - a[:, 2]) / 1e3
    ws = m[front[q]]
    wch = np.maximum(a[:, 10] - a[:, 2], 0) / 1e3
    asm = np.where(a[:, 11] > 0, a[:, 11] - a[:, 10], 0) / 1e3
    print(f"level {l:2d}: wait-children {wch.sum():8.0f} us, extend-add {asm.sum():8.0f} us (mean {asm.mean():5.1f})", end="  ")
    print(f"level {l:2d}: fronts {len(np.unique(front[q])):4d} tasks {q.sum():6d}  first start {us(a[:, 2].min()):8.1f}  last start {us(a[:, 2].max()):8.1f}  last end {us(a[:, 9].max()):8.1f}  CTA-time {dur.sum():9.0f} us  mean {dur.mean():6.1f} max {dur.max():7.1f}  mmax {ws.max()}")
# the widest front of every level: its chain
for l in np.unique(L):
    q = L == l
    fs = np.unique(front[q])
    f = fs[np.argmax(w[fs])]
    sel = front == f
    a = tr[sel]; ii = I[sel]; jj = J[sel]
    diag = (ii == jj) & (a[:, 8] > 0)
    ends = np.sort(a[diag][:, 9])
    line = f"level {l:2d} widest front {f} (m {m[f]}, w {w[f]}): start {us(a[:, 2].min()):8.1f} diag ends " + " ".join(f"{us(e):.0f}" for e in ends[:30]) + f"  all done {us(a[:, 9].max()):8.1f}"
    print(line)
    k0 = np.where((ii == 0) & (jj == 0))[0]
    if len(k0):
        d = a[k0[0]]
        print(f"          task (0,0): start {us(d[2]):8.1f} children done {us(d[10]):8.1f} assembled {us(d[11]):8.1f} ldlt {us(d[7]):8.1f} .. {us(d[8]):8.1f} end {us(d[9]):8.1f}")
    cbt = a[(a[:, 8] == 0) & (a[:, 4] == 0)]      # plain tasks without substitution / LDL^T: contribution block + ypre
    if len(cbt):
        print(f"          {len(cbt)} tail tasks: first start {us(cbt[:, 2].min()):8.1f}, updates done (mean) {us(cbt[:, 3].mean()):8.1f}, last end {us(cbt[:, 9].max()):8.1f}, mean duration {((cbt[:, 9] - cbt[:, 2]).mean()) / 1e3:6.1f} us, mean extend-add {((cbt[:, 11] - cbt[:, 10]).mean()) / 1e3:5.1f} us")
