import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
B = 8192
mod = pack_dense_models(range(B))
S = B200BatchNLS(B)
ptrs = S.upload(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
rec = S.solve_dev(ptrs, B); print("all", S.last_ms())
cost = rec[:, 2] * 30 + rec[:, 4] + rec[:, 5]
top = np.argsort(-cost)[:12]
for i in top: print(i, rec[i, :12])
S.close()
# time the top instances alone, and a random subset without them
for name, sel in (("top12", top), ("top1", top[:1]), ("rest-sample", np.setdiff1d(np.arange(2048), top))):
    m2 = {k: np.ascontiguousarray(v[sel]) for k, v in mod.items()}
    S = B200BatchNLS(len(sel))
    p = S.upload(m2["At"], m2["Bt"], m2["Ct"], m2["y"], m2["e"], m2["x0"])
    for _ in range(2):
        r = S.solve_dev(p, len(sel))
    print(name, len(sel), "ms", S.last_ms())
    S.close()
