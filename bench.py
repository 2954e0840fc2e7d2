#!/usr/bin/env python
"""bench.py -- KKT LDL^T factor+solve throughput of the B200 `linsolve` backend.

A *step* is the linear algebra of ONE Newton system of CaNNOLeS (reference/src/CaNNOLeS.jl
:1008-1052 with no inertia retry): `try_to_factorize` (COO->CSC accumulate, numeric LDL^T,
inertia counts) followed by `solve_ldl!`, on the KKT matrix of the named synthetic config at a
fixed iterate.  Default workload: C4 of BASELINE.json (2-D Poisson-constrained parameter
estimation on a 512^2 grid, N = 1 310 720 unknowns) -- the ~10^6-unknown system the north-star
target is quoted on.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3] [--size S]
  python bench.py --impl reference ...     # the reference's CPU path (restated; see below)

  value  factor+solve per second, whole job, inputs resident in HBM (b2_factorize_dev +
         b2_solve_dev), CUDA events on the engine's own stream, max over ranks
  e2e    the same through the reference-facing interface (B200Struct.try_to_factorize /
         solve_ldl) with pinned HOST buffers: H2D of vals and rhs, D2H of d and the inertia
         counts inside the timed region
  N > 1  one KKT system does not shard: every rank factors its own replica
         ("weak" scaling, no data-path collective); the C5 batch of independent instances is
         partitioned across ranks and reported in the "batched" sub-object.

`--impl reference`: Julia is not available offline, so the reference arm is the CPU
restatement of `cannoles(...; linsolve=:ldlfactorizations)`'s factor/solve path
(oracle/ldl_oracle.c: AMD + up-looking LDL^T, single-threaded like LDLFactorizations.jl).
"""
from __future__ import annotations

import argparse
import functools
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

EPS = 2.0 ** -52
METRIC = "kkt_factor_solve_per_s"
UNIT = "KKT factor+solve/s"
FALLBACK_HBM_GBS = 6650.0
# dram__bytes_read.sum + dram__bytes_write.sum of ONE factorization (all its kernels), from the ncu
# capture summarised in profiles/traffic_r02_i.txt; only valid for the default workload
KNOWN_TRAFFIC = {("c4", None, "nd"): 1424.0e6}
TRAFFIC_SOURCE = "profiles/traffic_r02_i.txt"


# ------------------------------------------------------------------------------------------
def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period=0.2):
        self.gpu, self.period = gpu_index, period
        self.samples, self._stop, self._th = [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); smax.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback"


# ------------------------------------------------------------------------------------------
def build_workload(args, ctor):
    from cannoles_b200.workloads import first_system, make_config
    nls, method, desc = make_config(args.workload, args.size)
    s, rhs = first_system(nls, method, ctor)
    return nls, method, desc, s, rhs


def config_dict(args, desc, st, extra=None):
    cfg = {"workload": f"{args.workload}: {desc}", "N": int(st["N"]), "nnz_coo": int(st["nnz"]),
           "nnzA": int(st["nnzA"]), "nnzL": int(st["nnzL"]), "flops_factor": float(st["flops"]),
           "l2_policy": ("factor panels + contribution blocks = %.0f MB per step, %s the 126 MB L2; no flush"
                         % (8e-6 * (st["nnzL_store"] + st["cb_store"]),
                            "larger than" if 8e-6 * (st["nnzL_store"] + st["cb_store"]) > 126 else
                            "NOT larger than (timing not L2-cold)")),
           "hessian_mode": None}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU factor/solve path (restated), rank 0 only."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import LDLFactStruct
    ncores = os.cpu_count() or 1
    ctor = functools.partial(LDLFactStruct)          # AMD ordering, as ldl_analyze does
    nls, method, desc, s, rhs = build_workload(args, ctor)
    O = s.LDLT
    # one probe repetition sizes the bounded sample
    t_probe, t3, ok, _ = O.time_factor_solve(s.vals, rhs, nls.nvar, EPS, reps=1)
    budget = 240.0
    K = max(1, min(args.steps, int(budget / max(t_probe, 1e-9))))
    W = min(args.warmup, 1) if t_probe > 5 else args.warmup
    for _ in range(max(0, W - 1)):               # the probe already was one warm-up
        O.time_factor_solve(s.vals, rhs, nls.nvar, EPS, reps=1)
    t0 = time.perf_counter()
    tt, t3, ok, d = O.time_factor_solve(s.vals, rhs, nls.nvar, EPS, reps=K)
    wall = time.perf_counter() - t0
    per = wall / K
    st = {"N": O.N, "nnz": len(s.vals), "nnzA": O.nnzA, "nnzL": O.nnzL, "flops": O.flops,
          "nnzL_store": O.nnzL, "cb_store": 0}
    sample = (f"{K} x full factor+solve of the same KKT system (requested steps={args.steps}; "
              f"capped so the run ends within ~{budget:.0f} s), AMD ordering, 1 thread "
              f"(LDLFactorizations.jl is single-threaded); host has {ncores} cores")
    val = 1.0 / per
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": per * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, desc, st, {"hessian_mode": method, "ordering": "amd"}),
            "phase_ms": {"assemble": t3[0] * 1e3, "factor": t3[1] * 1e3, "solve": t3[2] * 1e3},
            "inertia_ok": bool(ok),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "restated reference (Julia not available offline); MA57 excluded"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def measure_system(lib, args, workload, size, ordering_name, steps, warmup, local, dist, world, rank,
                   fp64_peak, hbm_peak, peak_src, with_cpu, sampler=None, preload_s=0.0):
    """value / e2e / roofline / parity gate of the factor+solve of ONE KKT system of a named config.
    Returns the fields of a bench line (dict)."""
    import ctypes as C
    from cannoles_b200 import _capi
    from cannoles_b200.linsolve import B200Struct
    from cannoles_b200.workloads import first_system, make_config
    ordering = {"nd": 0, "natural": 1, "amd": 3}[ordering_name]
    nls, method, desc = make_config(workload, size)

    def ctor(N, rows, cols, vals):
        return B200Struct(N, rows, cols, vals, ordering=ordering, device=local, refine_steps=args.refine,
                          shift_retries=False, nvar=nls.nvar, nequ=nls.nequ, ncon=nls.ncon)

    s, rhs = first_system(nls, method, ctor)
    B = s.LDLT
    N, nnz = B.N, len(s.vals)
    st = B.stats()
    d_host = np.zeros(N)
    B.register_host(rhs)
    B.register_host(d_host)
    h = B._h
    vp = C.c_void_p

    def chk(rc):
        if rc != 0:
            raise RuntimeError(_capi.last_error(lib))

    dv, dr, do = vp(), vp(), vp()
    chk(lib.b2_dev_malloc(C.byref(dv), nnz * 8))
    chk(lib.b2_dev_malloc(C.byref(dr), N * 8))
    chk(lib.b2_dev_malloc(C.byref(do), N * 8))
    chk(lib.b2_dev_upload(dv, s.vals.ctypes.data_as(vp), nnz * 8))
    chk(lib.b2_dev_upload(dr, rhs.ctypes.data_as(vp), N * 8))
    npos, nzero, nneg, brk = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
    ms5 = np.zeros(5)
    sweeps = [1]

    def step_dev():
        chk(lib.b2_factorize_dev(h, dv, EPS, C.byref(npos), C.byref(nzero), C.byref(nneg), C.byref(brk)))
        chk(lib.b2_last_timings(h, ms5.ctypes.data_as(_capi.pd)))
        t_asm, t_fac = ms5[1], ms5[2]
        chk(lib.b2_solve_dev(h, dr, do, 1, args.refine, None))
        chk(lib.b2_last_timings(h, ms5.ctypes.data_as(_capi.pd)))
        sweeps[0] = lib.b2_last_sweeps(h)
        return t_asm, t_fac, ms5[3]

    def step_host():
        ok = B.try_to_factorize(s.vals, nls.nvar, nls.nequ, nls.ncon, EPS)
        B.solve_ldl(rhs, d_host)
        return ok

    def barrier():
        chk(lib.b2_dev_sync())
        if dist is not None:
            dist.barrier()
        chk(lib.b2_dev_sync())

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also captures the CUDA graphs) and correctness gate ----------------------
    for _ in range(max(warmup, 3)):
        step_dev()
    ok = step_host()
    expected = (nls.nvar, 0, nls.nequ + nls.ncon, False)
    rho_used = 0.0
    # the reference's inertia correction (newton_system!, src/CaNNOLeS.jl:1029-1043): if the rho = 0
    # matrix does not factor with the expected inertia under this ordering (a Gauss-Newton KKT
    # matrix has a zero (1,1) block), raise rho as the reference does and time THAT system
    while (not ok or B.last_inertia != expected) and rho_used < 1e3:
        rho_used = EPS ** (1.0 / 3.0) if rho_used == 0.0 else 100.0 * rho_used
        s.vals[nnz - nls.nvar:] = rho_used
        ok = step_host()
    rho_reason = "inertia correction (rho = 0 fails the inertia test under this ordering)" if rho_used > 0 else None
    if ok and B.last_inertia == expected and rho_used == 0.0 and not (B.last_relres <= 1e-12):
        # A Gauss-Newton KKT matrix (zero (1,1) block) can pass the inertia test at rho = 0 under a
        # nested-dissection order that happens to eliminate every r-vertex before its x-neighbours and
        # still be numerically singular (C3: relres between 5e-3 and 3e-12 after refinement, depending on rounding --
        # above the 1e-12 acceptance bar either way).  The reference's AMD path
        # breaks down on that matrix and factors the rho0 system (tests/test_gpu_parity.py
        # ::test_c3_full_size_properties); time THAT system and say so.
        rho_used = EPS ** (1.0 / 3.0)
        s.vals[nnz - nls.nvar:] = rho_used
        ok = step_host()
        rho_reason = ("rho = 0 passes the inertia test under this ordering but the matrix is numerically singular "
                      "(relres > 1e-12 after refinement); the rho0 system the reference's AMD path factors is timed")
    inertia = B.last_inertia
    if not ok or inertia != expected:
        raise RuntimeError(f"wrong inertia {inertia}, expected {expected}")
    chk(lib.b2_dev_upload(dv, s.vals.ctypes.data_as(vp), nnz * 8))
    for _ in range(2):
        step_host()
    # untimed load before the timed region: the timed region of K steps is a fraction of a second,
    # shorter than nvidia-smi's sampling period -- the clock sampler starts here so that it sees
    # the same steps under load (the timed K steps follow without a gap)
    if sampler is not None and rank == 0:
        sampler.start()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < preload_s:
        step_dev()

    # ---- timed region 1: device-resident ---------------------------------------------------
    barrier()
    chk(lib.b2_timer_start(h))
    t0 = time.perf_counter()
    ph = np.zeros(3)
    for _ in range(steps):
        ph += step_dev()
    tms = C.c_double()
    chk(lib.b2_timer_stop(h, C.byref(tms)))
    barrier()
    wall_dev = time.perf_counter() - t0
    dev_ms = max(max_over_ranks(tms.value), 1e-9)
    ph = np.maximum(ph / steps, 1e-9)   # (1e-9 only ever bites on the CPU emulator used to debug this script)

    # ---- timed region 2: end to end through the reference-facing interface -----------------
    barrier()
    chk(lib.b2_timer_start(h))
    for _ in range(steps):
        step_host()
    chk(lib.b2_timer_stop(h, C.byref(tms)))
    barrier()
    e2e_ms = max(max_over_ranks(tms.value), 1e-9)
    clocks = sampler.stop() if (sampler is not None and rank == 0) else None
    relres = B.last_relres

    # ---- roofline --------------------------------------------------------------------------
    nnzA, nnzL = st["nnzA"], st["nnzL"]
    bytes_asm = 8 * nnz + 4 * nnz + 4 * nnzA + 8 * nnzA
    bytes_fact = 8 * (nnzA + nnzL + N)
    nsw = sweeps[0]                       # sweeps actually taken (adaptive refinement)
    nres = min(nsw, args.refine) if args.refine > 0 else 0
    bytes_solve = nsw * (2 * 8 * nnzL + 8 * 3 * N) + nres * (12 * nnzA + 16 * N)
    fact_tflops = st["flops"] / (ph[1] * 1e-3) / 1e12
    traffic = KNOWN_TRAFFIC.get((workload, size, ordering_name))
    roof = {"kernel": "numeric LDL^T factorization (one CUDA-graph launch: all fronts, all levels)",
            "bound": "tensor", "achieved": fact_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (fact_tflops / fp64_peak) if fp64_peak else None,
            "traffic": traffic,
            "traffic_note": ("bytes per factorization, ncu cold-cache sum over its kernels (%s); algorithmic bytes 8 (nnzA + nnzL + N) = %.1f MB"
                             % (TRAFFIC_SOURCE, bytes_fact / 1e6)),
            "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
            "algorithmic_flops_per_launch": st["flops"], "ms_per_launch": ph[1],
            "hbm_view": {"achieved": bytes_fact / (ph[1] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_fact / (ph[1] * 1e-3) / 1e9 / hbm_peak}}
    phases = {
        "assemble": {"bound": "hbm", "ms": ph[0], "achieved": bytes_asm / (ph[0] * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": bytes_asm / (ph[0] * 1e-3) / 1e9 / hbm_peak},
        "solve": {"bound": "hbm", "ms": ph[2], "achieved": bytes_solve / (ph[2] * 1e-3) / 1e9,
                  "peak": hbm_peak, "unit": "GB/s", "frac": bytes_solve / (ph[2] * 1e-3) / 1e9 / hbm_peak},
        "peak_source": peak_src}
    launches_step = int(st["launches_factor"] + st["launches_solve"] * nsw + 5 * nres)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample ------------
    cpu = None
    if with_cpu:
        from oracle import LDLFactStruct
        O = LDLFactStruct(N, s.rows, s.cols, s.vals)
        t, t3, okc, _ = O.time_factor_solve(s.vals, rhs, nls.nvar, EPS, reps=1)
        cpu = {"value": 1.0 / t, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "1 x full factor+solve of the same KKT system (AMD ordering, up-looking "
                         "LDL^T restated from LDLFactorizations.jl, 1 thread; host has %d cores)"
                         % (os.cpu_count() or 1),
               "phase_ms": {"assemble": t3[0] * 1e3, "factor": t3[1] * 1e3, "solve": t3[2] * 1e3},
               "inertia_ok": bool(okc)}
    for p in (dv, dr, do):
        lib.b2_dev_free(p)
    B.close()
    cfg_args = argparse.Namespace(workload=workload, size=size)
    return {"value": world * steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps,
            "config": config_dict(cfg_args, desc, st, {
                "hessian_mode": method, "ordering": ordering_name, "refine_steps_max": args.refine, "refine_tol": B.refine_tol,
                "solve_sweeps_used": nsw, "rho": rho_used, "rho_reason": rho_reason, "nsuper": int(st["nsuper"]), "nlevels": int(st["nlevels"]),
                "max_front": int(st["max_front"]), "parallelism": f"replicas x{world}"}),
            "phase_ms": {"assemble": ph[0], "factor": ph[1], "solve": ph[2]},
            "wall_ms_per_step": wall_dev / steps * 1e3,
            "inertia": list(inertia[:3]), "relres": relres,
            "roofline": roof, "roofline_phases": phases, "cpu_baseline": cpu,
            "e2e": {"value": world * steps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / steps,
                    "h2d_bytes_per_step": 8 * nnz + 8 * N, "d2h_bytes_per_step": 8 * N + 40},
            "gpu_launches": launches_step * steps, "clocks": clocks,
            "analyze_s": {"order": st["t_order"], "symbolic": st["t_symbolic"], "plan": st["t_plan"]}}


def cannoles_wall(lib, local, with_cpu=True):
    """BASELINE's second metric, "cannoles wall time": the restated reference loop
    (cannoles_b200/solver.py -- host Python/numpy model callbacks, CGLS, line search) with the
    B200 backend as `linsolve`, and beside it the same loop with the CPU oracle on the SAME
    elimination order (where that ends within seconds).  Equality of iter / nfact / nlinsolve and
    x to 1e-8 is the north-star acceptance."""
    import functools
    from cannoles_b200 import CaNNOLeSSolver, solve
    from cannoles_b200.linsolve import B200Struct
    from cannoles_b200.workloads import make_config
    out = []
    for cfg, size, cpu_too in (("c1", None, True), ("c2", None, True), ("c4", 256, True), ("c4", 512, False)):
        nls, method, desc = make_config(cfg, size)
        t0 = time.perf_counter()
        s = CaNNOLeSSolver(nls, linsolve=functools.partial(B200Struct, device=local, nvar=nls.nvar, nequ=nls.nequ,
                                                          ncon=nls.ncon), method=method)
        t_setup = time.perf_counter() - t0
        solve(s, nls, max_time=3600.0)          # warm: CUDA-graph capture, pinned buffers
        nls.reset_counters()
        t0 = time.perf_counter()
        st = solve(s, nls, max_time=3600.0)
        wall = time.perf_counter() - t0
        rec = {"config": f"{cfg}: {desc}", "N": nls.nvar + nls.nequ + nls.ncon,
               "b200": {"wall_s": wall, "setup_s": t_setup, "status": st.status, "iter": st.iter,
                        "nfact": st.solver_specific["nfact"], "nlinsolve": st.solver_specific["nlinsolve"],
                        "objective": st.objective, "primal_feas": st.primal_feas, "dual_feas": st.dual_feas}}
        perm = s.LDLT.perm
        xb = st.solution.copy()
        s.LDLT.close()
        if cpu_too and with_cpu:
            from oracle import LDLFactStruct
            nls2, _, _ = make_config(cfg, size)
            t0 = time.perf_counter()
            s2 = CaNNOLeSSolver(nls2, linsolve=functools.partial(LDLFactStruct, perm=perm), method=method)
            t_setup2 = time.perf_counter() - t0
            t0 = time.perf_counter()
            st2 = solve(s2, nls2, max_time=3600.0)
            wall2 = time.perf_counter() - t0
            rec["cpu_port"] = {"wall_s": wall2, "setup_s": t_setup2, "status": st2.status, "iter": st2.iter,
                               "nfact": st2.solver_specific["nfact"], "nlinsolve": st2.solver_specific["nlinsolve"],
                               "objective": st2.objective, "cores": 1,
                               "note": "same loop, oracle/ldl_oracle.c as linsolve on the B200 backend's elimination order"}
            rec["same_iter_nfact_nlinsolve"] = bool(
                (st.iter, st.solver_specific["nfact"], st.solver_specific["nlinsolve"]) ==
                (st2.iter, st2.solver_specific["nfact"], st2.solver_specific["nlinsolve"]))
            rec["x_rel_diff"] = float(np.linalg.norm(xb - st2.solution) / max(1.0, np.linalg.norm(st2.solution)))
            rec["speedup_wall"] = wall2 / wall
        out.append(rec)
    return out


def run_b200(args):
    import ctypes as C
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cannoles_b200 import _capi
    lib = _capi.load()
    if lib.b2_device_count() <= 0:
        raise RuntimeError("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    own, cub = C.c_double(), C.c_double()
    fp64_peak = None
    if hasattr(lib, "b2_measure_dgemm") and lib.b2_measure_dgemm(4096, 5, C.byref(own), C.byref(cub)) == 0:
        fp64_peak = cub.value
    with_cpu = rank == 0 and world == 1 and not args.no_cpu
    main = measure_system(lib, args, args.workload, args.size, args.ordering, args.steps, args.warmup, local, dist,
                          world, rank, fp64_peak, hbm_peak, peak_src, with_cpu, sampler=ClockSampler(local),
                          preload_s=args.preload)
    batched = None
    if not args.no_batched:
        batched = bench_batched(args, rank, world, local, dist, fp64_peak=fp64_peak, hbm_peak=hbm_peak)
    configs, wall = None, None
    if world == 1 and not args.no_configs:
        # the other single-system configs of BASELINE.json at their full sizes (parity-gated like the
        # main line: expected inertia, relres) -- fewer steps, they are reported, not the headline
        configs = {}
        for wl, order in (("c2", "nd"), ("c3", "nd")):
            try:
                r = measure_system(lib, args, wl, None, order, max(3, min(args.steps, 10)), 3, local, None, 1, 0,
                                   fp64_peak, hbm_peak, peak_src, with_cpu and wl == "c2")
                r.pop("clocks", None)
                configs[wl] = r
            except Exception as e:   # reported, never hidden
                configs[wl] = {"error": repr(e)}
        wall = cannoles_wall(lib, local, with_cpu=not args.no_cpu)
    if rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic"}
        for k in ("config", "phase_ms", "wall_ms_per_step", "inertia", "relres", "roofline", "roofline_phases",
                  "cpu_baseline", "e2e", "gpu_launches", "clocks", "analyze_s"):
            line[k] = main[k]
        if batched is not None:
            line["batched"] = batched
        if configs is not None:
            line["configs"] = configs
        if wall is not None:
            line["cannoles_wall"] = wall
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _cpu_nls_worker(i):
    """One instance of config 5 through the restated per-instance loop with the CPU oracle as
    `linsolve` (the cpu_baseline leg of the batched-NLS line; runs in a spawned worker process)."""
    from cannoles_b200.batched_nls import host_reference_loop
    from oracle import LDLFactStruct
    st, nls = host_reference_loop(i, LDLFactStruct)
    return (st.iter, st.solver_specific["nfact"], st.solver_specific["nlinsolve"])


def bench_nls(args, rank, world, local, dist, total, steps):
    """C5 for real: every instance of the batch runs the WHOLE CaNNOLeS iteration on the device
    (k_nls_dense: device prepare_newton_system!, rho retries, J'v products, CGLS, line search) to its
    final status.  A step = one batch of `total` instances partitioned over the ranks.
      latency     one batch alone (the slowest instance of the batch bounds it: a start that needs
                  ~1300 Newton systems keeps one SM busy for ~0.27 s while the others are done)
      value       `steps` batches in flight on the handle's lanes, model resident in HBM
      e2e         the same from pinned HOST model arrays to host records (H2D of the model, D2H of
                  the records inside the timed region, overlapped with the solves of other batches)
    One NCCL all_gather of the per-instance records at the end (the only collective)."""
    import ctypes as C
    from cannoles_b200 import _capi
    from cannoles_b200.batched import gather_records, partition
    from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
    lib = _capi.load()
    lo, hi = partition(total, rank, world)
    nb = hi - lo
    t0 = time.perf_counter()
    mod = pack_dense_models(range(lo, hi))
    t_gen = time.perf_counter() - t0
    S = B200BatchNLS(nb, device=local)
    names = ("At", "Bt", "Ct", "y", "e", "x0")
    arrs = [mod[k] for k in names]
    for a in arrs:
        S.kkt.register_host(a)
    ptrs = S.upload(*arrs)
    h = S.kkt._h
    # batches in flight: a rank's share of a batch shrinks with the number of ranks but the slowest
    # instance of the batch does not (its chain of ~1300 Newton systems takes ~0.1 s on one SM), so the
    # stream of batches has to be long enough for the other SMs to have work meanwhile
    K = min(128, max(steps, 16 * world))

    def chk(rc):
        if rc != 0:
            raise RuntimeError(_capi.last_error(lib))

    def sync():
        chk(lib.b2_dev_sync())
        if dist is not None:
            dist.barrier()
        chk(lib.b2_dev_sync())

    # latency of one batch alone: device-resident model, then host arrays (chunked upload)
    for _ in range(2):
        rec = S.solve_dev(ptrs, nb)
    lat_dev = S.last_ms()
    rec_h = S.solve(*arrs)
    lat_host = S.last_ms()
    assert np.array_equal(rec, rec_h)
    # throughput: K batches in flight
    recs = [np.zeros_like(rec) for _ in range(K)]
    for r in recs:
        S.kkt.register_host(r)
    drecs = []
    for _ in range(K):
        p = C.c_void_p()
        chk(lib.b2_dev_malloc(C.byref(p), rec.nbytes))
        drecs.append(p)
    ms = C.c_double()
    for k in range(K):                  # warm every lane the timed legs use (a lane allocates its buffers on first use)
        S.submit(arrs, recs[k])
    S.wait()
    sync()
    chk(lib.b2b_timer_start(h))
    for k in range(K):
        S.submit_dev(ptrs, nb, drecs[k])
    S.wait()
    chk(lib.b2b_timer_stop(h, C.byref(ms)))
    sync()
    dev_ms = max(ms.value, 1e-9)
    for r in recs:
        r[:] = 0
    sync()
    chk(lib.b2b_timer_start(h))
    for k in range(K):
        S.submit(arrs, recs[k])
    S.wait()
    chk(lib.b2b_timer_stop(h, C.byref(ms)))
    sync()
    e2e_ms = max(ms.value, 1e-9)
    same = all(np.array_equal(r, rec) for r in recs)
    for p in drecs:
        lib.b2_dev_free(p)
    S.close()
    t0 = time.perf_counter()
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_ms, lat_dev, lat_host], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, lat_dev, lat_host = (float(v) for v in t.tolist())
        allrec = gather_records(rec, dist, device="cuda")
    else:
        allrec = rec
    gather_s = time.perf_counter() - t0
    if rank != 0:
        return None
    nfact, nsolve, iters = allrec[:, 2].sum(), allrec[:, 3].sum(), allrec[:, 1].sum()
    slow = int(np.argmax(allrec[:, 2]))
    cpu = None
    if world == 1 and not args.no_cpu:
        import multiprocessing as mp
        ncores = os.cpu_count() or 1
        nsample = min(total, 16 * ncores)
        idx = [int(i) for i in np.linspace(0, total - 1, nsample).astype(int)]
        t0c = time.perf_counter()
        with mp.get_context("spawn").Pool(ncores) as pool:
            pool.map(_cpu_nls_worker, idx[:ncores])          # start-up (imports) outside the timing
            t0c = time.perf_counter()
            res = pool.map(_cpu_nls_worker, idx, chunksize=4)
            tc = time.perf_counter() - t0c
        same_counts = sum(1 for i, r in zip(idx, res) if r == tuple(int(v) for v in allrec[i, 1:4]))
        cpu = {"value": nsample / tc, "unit": "instances solved/s", "cores": ncores, "kind": "port",
               "sample": f"{nsample} of the {total} instances through the restated per-instance loop "
                         f"(cannoles_b200/solver.py: Python/numpy callbacks, CGLS, line search) with oracle/ldl_oracle.c "
                         f"as linsolve (AMD), {ncores} worker processes; Julia is not available offline",
               "same_iter_nfact_nlinsolve": f"{same_counts} of {nsample}"}
    return {"metric": "batched_nls_instances_solved_per_s", "unit": "instances solved/s",
            "value": total * K / (dev_ms * 1e-3), "ms_per_step": dev_ms / K,
            "e2e": {"value": total * K / (e2e_ms * 1e-3), "unit": "instances solved/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": int(sum(a.nbytes for a in arrs)) * world,
                    "d2h_bytes_per_step": int(rec.nbytes) * world,
                    "records_identical_to_the_single_batch_run": bool(same)},
            "kkt_factor_solves_per_s": float(nfact) * K / (dev_ms * 1e-3),
            "latency_ms": {"one_batch_device_resident": lat_dev, "one_batch_host_arrays": lat_host,
                           "note": "bounded by the slowest instance of the batch (instance %d: %d factorizations, %d line-search "
                                   "backtracks); batches in flight fill the SMs it leaves idle" % (slow, allrec[slow, 2], allrec[slow, 4])},
            "steps": K, "batches_in_flight": K, "n_gpus": world, "scaling": "strong",
            "config": {"workload": "c5: %d independent constrained NLS (n=64, m=128, 16 constraints), multi-start x0 ~ N(0,1), "
                                   "solved to the reference's default tolerances" % total,
                       "batch_total": total, "batch_per_rank": nb, "params": "ParamCaNNOLeS / solve! defaults (max_time = Inf)"},
            "totals": {"iter": float(iters), "nfact": float(nfact), "nlinsolve": float(nsolve),
                       "status_first_order": int((allrec[:, 0] == 1).sum()), "records_gathered": int(allrec.shape[0])},
            "gpu_launches": 2 * K, "gather_s": gather_s, "generate_s": t_gen, "cpu_baseline": cpu}


def bench_batched(args, rank, world, local, dist, total=None, steps=None, fp64_peak=None, hbm_peak=None):
    """C5: `total` independent instances partitioned over the ranks; a step = fused
    factorize + inertia + solve of every instance of the rank (values resident in HBM).
    Returns the "batched" sub-object of the bench line (rank 0) or None."""
    import ctypes as C
    from cannoles_b200 import _capi
    from cannoles_b200.batched import B200BatchStruct, gather_records, partition
    from cannoles_b200.linsolve import B200Error
    from cannoles_b200.workloads import dense_batch_systems
    lib = _capi.load()
    total = int(total or os.environ.get("B2_BENCH_BATCH", 8192))
    steps = int(steps or max(3, min(args.steps, 20)))
    lo, hi = partition(total, rank, world)
    nb = hi - lo
    t0 = time.perf_counter()
    s, vals, rhs = dense_batch_systems(range(lo, hi))
    t_gen = time.perf_counter() - t0
    nv, ne, nc = s.nvar, s.nequ, s.ncon
    N, nnz = nv + ne + nc, vals.shape[1]
    Bt = B200BatchStruct(N, s.rows, s.cols, nb, nv, ne, nc, device=local)
    st = Bt.stats()
    vp = C.c_void_p

    def chk(rc):
        if rc != 0:
            raise B200Error(_capi.last_error(lib))

    dv, dr, do, dc = vp(), vp(), vp(), vp()
    chk(lib.b2_dev_malloc(C.byref(dv), vals.nbytes))
    chk(lib.b2_dev_malloc(C.byref(dr), rhs.nbytes))
    chk(lib.b2_dev_malloc(C.byref(do), rhs.nbytes))
    chk(lib.b2_dev_malloc(C.byref(dc), nb * 32))
    chk(lib.b2_dev_upload(dv, vals.ctypes.data_as(vp), vals.nbytes))
    chk(lib.b2_dev_upload(dr, rhs.ctypes.data_as(vp), rhs.nbytes))
    h = Bt._h

    def sync():
        chk(lib.b2_dev_sync())
        if dist is not None:
            dist.barrier()
        chk(lib.b2_dev_sync())

    for _ in range(3):
        chk(lib.b2b_factor_solve_dev(h, dv, dr, do, None, EPS, 1, 0, dc))
    sync()
    chk(lib.b2b_timer_start(h))
    for _ in range(steps):
        chk(lib.b2b_factor_solve_dev(h, dv, dr, do, None, EPS, 1, 0, dc))
    ms = C.c_double()
    chk(lib.b2b_timer_stop(h, C.byref(ms)))
    sync()
    dev_ms = max(ms.value, 1e-9)
    # end to end from pinned-size host arrays through the public verb
    d = np.zeros((nb, N))
    for arr in (vals, rhs, d):
        Bt.register_host(arr)
    ok = Bt.factor_solve(vals, rhs, d)
    sync()
    t1 = time.perf_counter()
    chk(lib.b2b_timer_start(h))
    e2e_steps = max(2, steps // 4)
    for _ in range(e2e_steps):
        ok = Bt.factor_solve(vals, rhs, d)
    chk(lib.b2b_timer_stop(h, C.byref(ms)))
    sync()
    e2e_ms = max(ms.value, 1e-9)
    e2e_wall = (time.perf_counter() - t1) * 1e3
    # per-instance record: [ok, npos, nzero, nneg, ||d||]
    rec = np.stack([ok.astype(np.float64), Bt.npos.astype(np.float64), Bt.nzero.astype(np.float64),
                    Bt.nneg.astype(np.float64), np.linalg.norm(d, axis=1)], axis=1)
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0].item()), float(t[1].item())
        allrec = gather_records(rec, dist, device="cuda")
    else:
        allrec = rec
    for p in (dv, dr, do, dc):
        lib.b2_dev_free(p)
    Bt.close()
    nls = bench_nls(args, rank, world, local, dist, total, steps)
    if rank != 0:
        return None
    flops_inst = st["flops"]
    # CPU baseline for the batch: the oracle port on a bounded sample of the same instances, one
    # oracle handle per host thread (ctypes releases the GIL), all host cores
    cpu = None
    if world == 1 and not args.no_cpu:
        import threading
        from oracle import LDLFactStruct
        ncores = os.cpu_count() or 1
        nsample = min(nb, 32 * ncores)
        handles = [LDLFactStruct(N, s.rows, s.cols, vals[0].copy()) for _ in range(ncores)]

        def work(tix):
            O = handles[tix]
            dd = np.zeros(N)
            for bidx in range(tix, nsample, ncores):
                O.try_to_factorize(vals[bidx], nv, ne, nc, EPS)   # (a wrong inertia at rho = 0 is legitimate)
                O.solve_ldl(rhs[bidx], dd)

        th = [threading.Thread(target=work, args=(i,)) for i in range(ncores)]
        t0c = time.perf_counter()
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        tc = time.perf_counter() - t0c
        cpu = {"value": nsample / tc, "unit": "instances/s", "cores": ncores, "kind": "port",
               "sample": f"{nsample} of the {total} instances, factor+solve each (AMD, up-looking LDL^T), "
                         f"{ncores} host threads"}
    out = {"metric": "batched_kkt_factor_solve_per_s", "unit": "instances/s",
           "value": total * steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps,
           "e2e": {"value": total * e2e_steps / (e2e_ms * 1e-3), "unit": "instances/s",
                   "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": e2e_wall / e2e_steps,
                   "h2d_bytes_per_step": int(vals.nbytes + 2 * rhs.nbytes) * world,
                   "d2h_bytes_per_step": int(rhs.nbytes + nb * 32) * world},
           "config": {"workload": "c5: %d independent constrained NLS (n=%d, m=%d, %d constraints), "
                                  "first Newton system of each" % (total, nv, ne, nc),
                      "N": N, "nnz_coo": nnz, "batch_total": total, "batch_per_rank": nb,
                      "nnzL": int(st["nnzL"]), "flops_factor_per_instance": flops_inst,
                      "nsuper": int(st["nsuper"]), "nlevels": int(st["nlevels"])},
           "steps": steps, "n_gpus": world, "scaling": "strong",
           "inertia_ok_at_rho0": int((allrec[:, 0] == 1).sum()), "records_gathered": int(allrec.shape[0]),
           "note_inertia": "instances whose exact-Hessian KKT matrix has the wrong inertia at rho = 0 are the ones newton_system! retries with rho > 0; they are factorized and counted but not solved in this step",
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "kernel": "k_batched<256> (one launch = the rank's instances)",
                        "achieved": flops_inst * (total / world) * steps / (dev_ms * 1e-3) / 1e12,
                        "peak": fp64_peak,
                        "frac": (flops_inst * (total / world) * steps / (dev_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                        "note": "per GPU; algorithmic flops sum_j(c_j^2+3c_j) per instance; peak = cuBLAS DGEMM measured in this run",
                        "hbm_view": {"achieved": (8.0 * nnz + 16.0 * N) * (total / world) * steps / (dev_ms * 1e-3) / 1e9,
                                     "peak": hbm_peak, "unit": "GB/s",
                                     "frac": ((8.0 * nnz + 16.0 * N) * (total / world) * steps / (dev_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None}},
           "cpu_baseline": cpu, "generate_s": t_gen}
    if nls is not None:
        out["nls"] = nls
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4"])
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--ordering", default="nd", choices=["nd", "amd", "natural"])
    ap.add_argument("--refine", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 lines and the cannoles wall-time runs")
    ap.add_argument("--preload", type=float, default=1.5,
                    help="seconds of untimed steps right before the timed region (the clock sampler runs from there)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
