import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200 import _capi
lib = _capi.bind_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb2_timing.so"))
_capi._LIB = lib
from cannoles_b200.linsolve import B200Struct
from tests.problems import random_kkt
EPS = 2.0 ** -52
nv, ne, nc = 250, 350, 100
N, r, c, v = random_kkt(nv, ne, nc, 0.6, 33)
B = B200Struct(N, r, c, v, nvar=nv, nequ=ne, ncon=nc, ordering=1, shift_retries=False)
for _ in range(3):
    ok = B.try_to_factorize(v, nv, ne, nc, EPS)
print(ok, B.stats()["max_front"], B.timings())
d = np.zeros(N)
for _ in range(3):
    B.solve_ldl(np.ones(N), d)
print("solve", B.timings(), B.last_relres)
out = (C.c_longlong * 64)()
lib.b2_debug_clocks.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.b2_debug_clocks(out))
t = list(out)
print("load S      ", t[1] - t[0])
print("ldlt total  ", t[2] - t[1], " panels", t[10], " trailing", t[11])
print("Lr build    ", t[3] - t[2])
print("R load      ", t[4] - t[3])
print("substitution", t[5] - t[4])
print("bwd_big CTA0 (last block of the front): init+gather", t[55]-t[54], " rows below", t[56]-t[55], " later blocks", t[57]-t[56], " substitution", t[58]-t[57], " publish", t[59]-t[58])
print("update: first load", t[21]-t[20], " main loop", t[22]-t[21], " epilogue", t[23]-t[22])
