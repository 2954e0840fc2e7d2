"""-m "not gpu": the device-resident batched CaNNOLeS loop (csrc/nls_kernels.cuh) on the CPU
emulator against the restated per-instance loop driven by the oracle on the same elimination
order: status, iter, nfact, nlinsolve, nbk, evaluation counts equal; x / objective / ||c|| to 1e-8;
the COO values of the first Newton system (device prepare_newton_system!) against the host fill."""
import ctypes as C

import numpy as np
import pytest

from cannoles_b200 import _capi
from tests import nls_checks as nc


def test_param_struct_matches_header(emu_lib):
    p = _capi.NLSParams()
    emu_lib.b2_nls_default_params(C.byref(p))
    eps = 2.0 ** -52
    assert C.sizeof(_capi.NLSParams) == 15 * 8 + 6 * 4
    assert (p.eig_tol, p.rho0, p.kappa_largeinc, p.max_eval, p.max_inner) == (eps, eps ** (1 / 3), 100.0, 100000, 10000)
    assert p.rho_max == eps ** -2 and p.delta_min == eps ** 0.5 and p.gamma_A == eps ** 0.25


@pytest.mark.parametrize("scale", [1.0, 10.0, 30.0])
def test_small_instances_match_host_loop(emu_lib, oracle_cls, scale):
    """scale 1: extrapolation steps only; 10 / 30: rho retries (nfact > nlinsolve) and line searches
    with backtracking (nbk > 0) -- every branch of the inner loop."""
    rec, out = nc.check_batch_nls(emu_lib, oracle_cls, range(4), 12, 20, 4, x0_scale=scale)
    if scale > 1:
        assert any(o["nfact"] > o["nlinsolve"] for o in out)
        assert any(o["nbk"] > 0 for o in out)


def test_host_verb_chunked(emu_lib, oracle_cls):
    """The host-buffer verb (chunked upload, one kernel per chunk) gives the records of the device verb."""
    rec_a, _ = nc.check_batch_nls(emu_lib, oracle_cls, range(5), 10, 16, 3, use_host_verb=True, chunk=2)
    rec_b, _ = nc.check_batch_nls(emu_lib, oracle_cls, range(5), 10, 16, 3)
    assert np.array_equal(rec_a, rec_b)


def test_async_submissions(emu_lib, oracle_cls):
    """b2b_nls_dense_submit x 3 + b2b_nls_wait == the synchronous verb."""
    from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
    mod = pack_dense_models(range(4), 10, 16, 3)
    arrs = [mod[k] for k in ("At", "Bt", "Ct", "y", "e", "x0")]
    S = B200BatchNLS(4, 10, 16, 3, _lib=emu_lib)
    try:
        ref = S.solve(*arrs)
        recs = [np.zeros_like(ref) for _ in range(3)]
        for r in recs:
            S.submit(arrs, r)
        S.wait()
    finally:
        S.close()
    assert all(np.array_equal(r, ref) for r in recs)


def test_unconstrained_instances(emu_lib, oracle_cls):
    nc.check_batch_nls(emu_lib, oracle_cls, range(2), 8, 14, 0)


def test_full_size_instance(emu_lib, oracle_cls):
    nc.check_batch_nls(emu_lib, oracle_cls, [3], 64, 128, 16)


def test_limits_are_honoured(emu_lib, oracle_cls):
    """max_iter / max_inner / max_eval end the loop with the reference's status."""
    nc.check_batch_nls(emu_lib, oracle_cls, range(2), 12, 20, 4, x0_scale=10.0, params_kw={"max_iter": 2},
                       loop_kw={"max_iter": 2})
    nc.check_batch_nls(emu_lib, oracle_cls, [3], 12, 20, 4, x0_scale=30.0, params_kw={"max_eval": 12},
                       loop_kw={"max_eval": 12})
    nc.check_batch_nls(emu_lib, oracle_cls, [3], 12, 20, 4, x0_scale=30.0, params_kw={"max_inner": 1},
                       loop_kw={"max_inner": 1})


def test_wrong_layout_is_rejected(emu_lib):
    from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
    from cannoles_b200.linsolve import B200Error
    S = B200BatchNLS(2, 8, 12, 2, _lib=emu_lib)
    try:
        mod = pack_dense_models(range(2), 8, 12, 2)
        S.m = 11   # model dimensions that are not the analysed KKT layout
        with pytest.raises(B200Error, match="differ from the analysed"):
            S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
    finally:
        S.close()


def test_partitioned_batch_gloo_world2(tmp_path, emu_lib):
    """SURVEY 8(e) for the batched loop: two ranks (gloo, CPU emulator) each solve their contiguous block
    of the instances, one all_gather of the records at the end == the records of a single-rank run."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "import torch.distributed as dist\n"
        "from cannoles_b200 import _capi\n"
        "from cannoles_b200.batched import partition, gather_records\n"
        "from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models\n"
        f"lib = _capi.bind_library({os.path.join(root, 'tests', 'hostsim', 'libb2_emu.so')!r})\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "total, dims = 5, (10, 16, 3)\n"
        "def run(lo, hi):\n"
        "    mod = pack_dense_models(range(lo, hi), *dims)\n"
        "    S = B200BatchNLS(hi - lo, *dims, _lib=lib)\n"
        "    rec = S.solve(mod['At'], mod['Bt'], mod['Ct'], mod['y'], mod['e'], mod['x0'])\n"
        "    S.close()\n"
        "    return rec\n"
        "lo, hi = partition(total, r, w)\n"
        "allrec = gather_records(run(lo, hi), dist)\n"
        "assert allrec.shape[0] == total\n"
        "assert np.array_equal(allrec, run(0, total))\n"
        "assert (allrec[:, 0] == 1).all()\n"
        "dist.barrier(); dist.destroy_process_group()\n"
        "print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_nan_start_is_reported_like_the_reference_error(emu_lib):
    """`Initial point gives Inf or Nan` (src/CaNNOLeS.jl:484-487) is an error in the reference; the
    device loop reports it as status 8 for that instance and solves the others."""
    from cannoles_b200.batched_nls import B200BatchNLS, STATUS, pack_dense_models
    mod = pack_dense_models(range(3), 10, 16, 3)
    mod["x0"][1, 2] = np.nan
    S = B200BatchNLS(3, 10, 16, 3, _lib=emu_lib)
    try:
        rec = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
    finally:
        S.close()
    assert [STATUS[int(s)] for s in rec[:, 0]] == ["first_order", "error: Initial point gives Inf or Nan", "first_order"]
    assert rec[1, 2] == 0 and rec[1, 3] == 0          # no factorization, no solve
