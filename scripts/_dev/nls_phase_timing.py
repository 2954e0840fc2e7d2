"""In-kernel phase clocks of k_nls_dense (block 0), -DB2_TIMING build (scripts/_dev/build_timing.sh)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200 import _capi
lib = _capi.bind_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb2_timing.so"))
_capi._LIB = lib
from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 4
mod = pack_dense_models(range(B))
S = B200BatchNLS(B)
ptrs = S.upload(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
lib.b2b_debug_clocks.argtypes = [C.POINTER(C.c_longlong)]
out = (C.c_longlong * 64)()
for _ in range(2):
    rec = S.solve_dev(ptrs, B)
    lib.b2b_debug_clocks(out)
t = list(out)
print("kernel %.3f ms; block 0: instances %d, systems %d" % (S.last_ms(), t[5], t[3]))
ns = max(t[3], 1)
print("per instance total %d cycles" % (t[0] / max(t[5], 1)))
print("per system: fill_vals %d, assemble+factor+solve %d, step/line search evals %d, trial products+norms+accept %d" % (t[1] / ns, t[2] / ns, t[6] / ns, t[4] / ns))
print("per instance: cgls %d" % (t[7] / max(t[5], 1)))
print("totals: nfact %d nbk %d neval %d iter %d" % (rec[:,2].sum(), rec[:,4].sum(), rec[:,5].sum(), rec[:,1].sum()))
print("last batched_instance: assemble", t[31]-t[30], "factor", t[40]-t[31], "inertia", t[41]-t[40], "solve", t[42]-t[41])
S.close()
