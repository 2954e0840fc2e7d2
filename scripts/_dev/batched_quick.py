import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200.batched import B200BatchStruct
from cannoles_b200.workloads import dense_batch_systems
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
s, vals, rhs = dense_batch_systems(range(nb))
print("built", flush=True)
Bt = B200BatchStruct(208, s.rows, s.cols, nb, 64, 128, 16)
print("analyzed", flush=True)
d = np.zeros((nb, 208))
t = time.time()
ok = Bt.factor_solve(vals, rhs, d)
print("done", ok.sum(), Bt.last_ms(), time.time() - t, flush=True)
