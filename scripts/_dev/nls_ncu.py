"""Two launches of k_nls_dense on 296 instances of config 5 (for `ncu --kernel-name regex:k_nls_dense --launch-skip 1 -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
B = 296
mod = pack_dense_models(range(B))
S = B200BatchNLS(B)
ptrs = S.upload(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
for _ in range(2):
    rec = S.solve_dev(ptrs, B)
print("ms", S.last_ms(), "nfact", rec[:, 2].sum())
S.close()
