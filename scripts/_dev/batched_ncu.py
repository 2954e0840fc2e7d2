import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cannoles_b200.batched import B200BatchStruct
from cannoles_b200.workloads import dense_batch_systems
nb = 296
s, vals, rhs = dense_batch_systems(range(nb))
Bt = B200BatchStruct(208, s.rows, s.cols, nb, 64, 128, 16)
d = np.zeros((nb, 208))
for _ in range(3):
    ok = Bt.factor_solve(vals, rhs, d)
print(ok.sum(), Bt.last_ms())
