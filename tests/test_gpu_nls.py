"""-m gpu: the device-resident batched CaNNOLeS loop (k_nls_dense) on the B200 through the C ABI,
against the restated per-instance loop driven by the oracle on the same elimination order
(status, iter, nfact, nlinsolve, nbk, evaluation counts equal; x / objective / ||c|| to 1e-8), and
at the BASELINE batch size (8192) through size-independent properties."""
import numpy as np
import pytest

from tests import nls_checks as nc

pytestmark = pytest.mark.gpu


def test_sample_of_c5_matches_host_loop(gpu_lib, oracle_cls):
    """256 instances spread over the 8192 of config 5 (seeds 1000 + i)."""
    nc.check_batch_nls(gpu_lib, oracle_cls, range(0, 8192, 32))


def test_hard_starts_match_host_loop(gpu_lib, oracle_cls):
    """x0 scaled by 10: rho retries and backtracking line searches in most instances."""
    rec, out = nc.check_batch_nls(gpu_lib, oracle_cls, range(48), x0_scale=10.0)
    assert sum(o["nfact"] > o["nlinsolve"] for o in out) >= 8
    assert any(o["nbk"] > 0 for o in out)


def test_host_verb_chunks_and_determinism(gpu_lib, oracle_cls):
    """Host-buffer verb (chunked upload on a copy stream, three chunk kernels in flight) == device
    verb == itself run twice, bit for bit."""
    from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
    inst = range(100, 700)
    mod = pack_dense_models(inst)
    S = B200BatchNLS(len(inst), _lib=gpu_lib)
    try:
        a = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"], chunk=150)
        b = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"], chunk=64)
        ptrs = S.upload(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
        c = S.solve_dev(ptrs, len(inst))
        arrs = [mod[k] for k in ("At", "Bt", "Ct", "y", "e", "x0")]
        for x in arrs:
            S.kkt.register_host(x)
        recs = [np.zeros_like(a) for _ in range(30)]     # more submissions than lanes: lanes are reused in stream order
        for r in recs:
            S.kkt.register_host(r)
            S.submit(arrs, r)
        S.wait()
    finally:
        S.close()
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert all(np.array_equal(a, r) for r in recs)
    assert (a[:, 0] == 1).all()


def test_shared_model_multistart(gpu_lib, oracle_cls):
    """One model, many starts (shared_model = 1): every start that converges reaches a first-order
    point of the same problem; starts equal to an instance's own x0 reproduce that instance."""
    from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
    mod = pack_dense_models([7])
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal((64, 64))
    x0[0] = mod["x0"][0]
    S = B200BatchNLS(64, _lib=gpu_lib)
    try:
        rec = S.solve(mod["At"][0], mod["Bt"][0], mod["Ct"][0], mod["y"][0], mod["e"][0], x0, shared_model=True)
        one = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
    finally:
        S.close()
    assert np.array_equal(rec[0], one[0])
    assert (rec[:, 0] == 1).all()
    assert (rec[:, 8] <= 1e-6).all()          # ||c(x)||


def test_full_batch_properties(gpu_lib, oracle_cls):
    """All 8192 instances of config 5: every instance ends first_order with a feasible point, the
    counters are consistent, and a second run is bit-identical."""
    from cannoles_b200.batched_nls import B200BatchNLS, pack_dense_models
    B = 8192
    mod = pack_dense_models(range(B))
    S = B200BatchNLS(B, _lib=gpu_lib)
    try:
        for a in mod.values():
            S.kkt.register_host(a)
        rec = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"])
        rec2 = S.solve(mod["At"], mod["Bt"], mod["Ct"], mod["y"], mod["e"], mod["x0"], chunk=1000)
    finally:
        S.close()
    assert np.array_equal(rec, rec2)
    assert (rec[:, 0] == 1).all(), np.unique(rec[:, 0], return_counts=True)
    assert (rec[:, 2] >= rec[:, 3]).all() and (rec[:, 3] >= rec[:, 1]).all()     # nfact >= nlinsolve >= iter
    assert (rec[:, 5] == rec[:, 6]).all()                                         # residual and constraint evaluations pair up
    assert (rec[:, 8] <= 1e-6).all() and np.isfinite(rec).all()
    # a handful against the host loop, taken from the full-batch run (other instances around them)
    perm = None
    from cannoles_b200.batched_nls import host_reference_loop
    import functools
    S2 = B200BatchNLS(1, _lib=gpu_lib)
    perm = S2.kkt.perm
    S2.close()
    for i in (0, 4095, 8191):
        st, nls = host_reference_loop(i, functools.partial(oracle_cls, perm=perm))
        nc.compare_instance(rec[i], st, nls, 64, 16, tag=f"instance {i}")
