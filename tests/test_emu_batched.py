"""-m "not gpu": the batched kernel (csrc/batched_kernels.cuh) on the CPU emulator against the
oracle, the masked rho driver, the rank partition and the gloo gather (world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cannoles_b200.batched import partition
from tests import batched_checks as bc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ordering", [0, 3])
def test_batched_against_oracle(emu_lib, oracle_cls, ordering):
    bc.check_batch_against_oracle(emu_lib, oracle_cls, 20, 26, 6, 0.15, 51, batch=3, ordering=ordering)


def test_batched_unconstrained_and_tiny(emu_lib, oracle_cls):
    bc.check_batch_against_oracle(emu_lib, oracle_cls, 9, 14, 0, 0.3, 52, batch=2)
    bc.check_batch_against_oracle(emu_lib, oracle_cls, 2, 2, 1, 1.0, 53, batch=2, ordering=1)


def test_batched_newton_system_matches_per_instance_driver(emu_lib, oracle_cls):
    bc.check_batched_newton_system(emu_lib, oracle_cls)


def test_too_large_for_shared_memory_is_rejected(emu_lib):
    from cannoles_b200.batched import B200BatchStruct
    from cannoles_b200.linsolve import B200Error
    n = 300
    k = np.arange(1, n + 1, dtype=np.int64)
    with pytest.raises(B200Error, match="shared memory"):
        B200BatchStruct(n, k, k, 2, n, 0, 0, _lib=emu_lib)


def test_partition_is_contiguous_and_complete():
    for B in (8192, 10, 7):
        for W in (1, 2, 4, 8):
            blocks = [partition(B, r, W) for r in range(W)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_gather_records_gloo_world2(tmp_path):
    """The end-of-run gather of per-instance records over 2 ranks (gloo, CPU)."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch.distributed as dist\n"
        "from cannoles_b200.batched import partition, gather_records\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = partition(11, r, w)\n"
        "rec = np.stack([np.arange(lo, hi, dtype=np.float64), np.full(hi - lo, float(r))], axis=1)\n"
        "allrec = gather_records(rec, dist)\n"
        "assert allrec.shape == (11, 2), allrec.shape\n"
        "assert np.array_equal(allrec[:, 0], np.arange(11.0))\n"
        "assert allrec[0, 1] == 0 and allrec[-1, 1] == 1\n"
        "dist.barrier(); dist.destroy_process_group()\n"
        "print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
