"""Device / end-to-end time of the batched CaNNOLeS loop on B instances of config 5."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t = time.time(); mod = pack_dense_models(range(B)); mod["x0"] *= scale; print("generate %.2f s" % (time.time() - t))
S = B200BatchNLS(B)
for a in mod.values(): S.kkt.register_host(a)
ptrs = S.upload(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
for _ in range(3):
    rec = S.solve_dev(ptrs, B); ms = S.last_ms()
    print("dev  %.3f ms  %.0f inst/s  nfact %d nlinsolve %d iter %d  -> %.0f KKT factor+solve/s  %.3f us/factor/SM" % (
        ms, B / ms * 1e3, rec[:, 2].sum(), rec[:, 3].sum(), rec[:, 1].sum(), rec[:, 2].sum() / ms * 1e3, ms * 1e3 * 148 / rec[:, 2].sum()))
for ch in (0, 256, 512, 1024):
    t = time.time(); rec = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"], chunk=ch); w = time.time() - t
    print("host chunk %4d  %.3f ms (wall %.3f)  %.0f inst/s" % (ch, S.last_ms(), w * 1e3, B / S.last_ms() * 1e3))
import ctypes as C
lib = S._lib
arrs = [mod[k] for k in ("At", "Bt", "Ct", "y", "e", "x0")]
for K in (4, 12, 24):
    recs = [np.zeros_like(rec) for _ in range(K)]
    for r in recs: S.kkt.register_host(r)
    for leg in ("host", "dev"):
        ms = C.c_double()
        if leg == "dev":
            drecs = []
            for _ in range(K):
                p = C.c_void_p(); lib.b2_dev_malloc(C.byref(p), rec.nbytes); drecs.append(p)
        lib.b2_dev_sync()
        lib.b2b_timer_start(S.kkt._h)
        t = time.time()
        for k in range(K):
            if leg == "host": S.submit(arrs, recs[k])
            else: S.submit_dev(ptrs, B, drecs[k])
        S.wait()
        lib.b2b_timer_stop(S.kkt._h, C.byref(ms))
        print("pipelined %s K=%d: %.3f ms/step (wall %.3f)  %.0f inst/s" % (leg, K, ms.value / K, (time.time() - t) * 1e3 / K, B * K / ms.value * 1e3))
        if leg == "dev":
            for p in drecs: lib.b2_dev_free(p)
    assert all(np.array_equal(r, rec) for r in recs)
print("status", np.unique(rec[:, 0], return_counts=True), "iter", np.unique(rec[:, 1], return_counts=True))
S.close()
