"""First-contact GPU script: parity against the oracle on small instances, then phase timings
of the single-system engine on configs C2 / C4.  Writes gpurun_out/gpu_check.jsonl."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cannoles_b200.linsolve import B200Struct  # noqa: E402
from cannoles_b200.models import ExtRosenbrockLinEq, PoissonParamEst  # noqa: E402

EPS = 2.0 ** -52
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, "gpu_check.jsonl"), "a")


def emit(**kw):
    s = json.dumps(kw)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


from cannoles_b200.workloads import first_system  # noqa: E402,F401


def run(nls, method, ordering, tag, parity, reps=5):
    import functools
    t0 = time.time()
    kw = {}
    if os.environ.get("B2_CHECK_EMU"):   # debugging this script on a CPU-only box
        from cannoles_b200 import _capi
        kw["_lib"] = _capi.bind_library(os.path.join(os.path.dirname(OUT), "tests", "hostsim", "libb2_emu.so"))
    ctor = functools.partial(B200Struct, ordering=ordering, nvar=nls.nvar, nequ=nls.nequ,
                             ncon=nls.ncon, refine_steps=1, **kw)
    s, rhs = first_system(nls, method, ctor)
    B = s.LDLT
    t_an = time.time() - t0
    st = B.stats()
    vals = s.vals
    N = B.N
    d = np.zeros(N)
    rec = dict(tag=tag, ordering=ordering, N=N, nnz=len(vals), t_analyze=t_an,
               **{k: st[k] for k in ("nnzA", "nnzL", "nnzL_store", "cb_store", "nsuper", "nlevels",
                                     "max_front", "n_small", "n_large", "launches_factor",
                                     "launches_solve", "flops", "flops_store", "t_order",
                                     "t_symbolic", "t_plan", "bytes_device")})
    oks = []
    tf, ts, ta, tu = [], [], [], []
    ok = B.try_to_factorize(vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    rec["inertia_rho0"] = B.last_inertia
    rho = 0.0
    while not ok and rho < 1e10:   # inertia correction through the device-side shift path
        rho = EPS ** (1 / 3) if rho == 0 else rho * 100
        vals[-nls.nvar:] = rho
        ok = B.try_to_factorize(vals, nls.nvar, nls.nequ, nls.ncon, EPS)
    rec["rho_needed"] = rho
    rec["n_shift"] = B.n_shift
    rec["ms_factor_shift"] = B.timings()["factor"]
    B.shift_retries = False        # timed repetitions: full upload + assemble + factor
    for it in range(reps):
        t1 = time.time()
        ok = B.try_to_factorize(vals, nls.nvar, nls.nequ, nls.ncon, EPS)
        t2 = time.time()
        tm = B.timings()
        oks.append(bool(ok))
        t3 = time.time()
        B.solve_ldl(rhs, d)
        t4 = time.time()
        tm2 = B.timings()
        tu.append(tm["upload"]); ta.append(tm["assemble"]); tf.append(tm["factor"]); ts.append(tm2["solve"])
        rec.setdefault("wall_factor_ms", []).append((t2 - t1) * 1e3)
        rec.setdefault("wall_solve_ms", []).append((t4 - t3) * 1e3)
    rec.update(ok=oks, inertia=B.last_inertia, relres=B.last_relres, ms_upload=tu, ms_assemble=ta,
               ms_factor=tf, ms_solve=ts)
    if min(tf) > 0:
        rec["factor_gflops"] = st["flops"] / (min(tf) * 1e-3) / 1e9
    if parity:
        from oracle import LDLFactStruct
        O = LDLFactStruct(N, s.rows, s.cols, vals, perm=B.perm)
        ok2 = O.try_to_factorize(vals, nls.nvar, nls.nequ, nls.ncon, EPS)
        xo = np.zeros(N)
        if ok2:
            O.solve_ldl(rhs, xo)
        dB, dO = B.factor.d, O.factor.d
        rec.update(oracle_ok=bool(ok2), oracle_inertia=O.inertia(EPS),
                   nzval_bitexact=bool(np.array_equal(B.nzval, O.nzval)),
                   maxrel_dD=float(np.max(np.abs(dB - dO) / np.abs(dO))) if ok2 else None,
                   rel_dx=float(np.linalg.norm(d - xo) / np.linalg.norm(xo)) if ok2 else None)
    emit(**rec)
    B.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["parity", "c2", "c4s", "c4"]
    if "parity" in which:
        run(ExtRosenbrockLinEq(2000), "Newton_noFHess", 0, "c2-2000", True)
        run(PoissonParamEst(32), "Newton", 0, "c4-32", True)
        run(PoissonParamEst(64), "Newton", 0, "c4-64", True)
        run(PoissonParamEst(64), "Newton", 3, "c4-64-amd", True)
    if "san" in which:   # small cases for compute-sanitizer
        run(ExtRosenbrockLinEq(400), "Newton_noFHess", 0, "san-c2-400", True, reps=1)
        run(PoissonParamEst(16), "Newton", 0, "san-c4-16", True, reps=1)
    if "c2" in which:
        run(ExtRosenbrockLinEq(100_000), "Newton_noFHess", 0, "c2", False)
    if "c4s" in which:
        run(PoissonParamEst(128), "Newton", 0, "c4-128", True)
        run(PoissonParamEst(256), "Newton", 3, "c4-256-amd", False)
    if "c4" in which:
        run(PoissonParamEst(512), "Newton", 3, "c4-512-amd", False)
        run(PoissonParamEst(512), "Newton", 0, "c4-512-nd", False)
