import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200 import _capi
lib = _capi.bind_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb2_timing.so"))
_capi._LIB = lib
from cannoles_b200.batched import B200BatchStruct
from cannoles_b200.workloads import dense_batch_systems
nb = 296
s, vals, rhs = dense_batch_systems(range(nb))
Bt = B200BatchStruct(208, s.rows, s.cols, nb, 64, 128, 16)
d = np.zeros((nb, 208))
for _ in range(3):
    ok = Bt.factor_solve(vals, rhs, d)
print(ok.sum(), Bt.last_ms(), Bt.stats()["nsuper"], Bt.stats()["nlevels"])
out = (C.c_longlong * 64)()
lib.b2b_debug_clocks.argtypes = [C.POINTER(C.c_longlong)]
lib.b2b_debug_clocks(out)
t = list(out)
print("assemble", t[31]-t[30], " (zero", t[20]-t[30], " single-slot entries", t[21]-t[20], " multi-slot entries", t[31]-t[21], ")")
print("ph0 (a)", t[32]-t[31], "(b)", t[33]-t[32])
print("ph1 (a)", t[34]-t[33], "(b)", t[35]-t[34])
print("root: (a) update", t[50], " dense LDL^T (cta_ldlt_packed)", t[51])
print("inertia", t[41]-t[40], "solve", t[42]-t[41], "total", t[42]-t[30])
print("solve split: load rhs + forward", t[43]-t[41], " diagonal", t[44]-t[43], " backward", t[45]-t[44], " (of which after the leaves pass of level 1:", t[45]-t[47], ", of level 0:", t[45]-t[46], ") store", t[42]-t[45])
