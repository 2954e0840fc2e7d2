import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        from cannoles_b200 import _capi
        return _capi.load().b2_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_cls():
    """The restated LDLFactStruct (CPU oracle)."""
    from oracle import LDLFactStruct
    return LDLFactStruct


@pytest.fixture(scope="session")
def emu_lib():
    """The product kernel sources compiled for the CPU emulator (tests/hostsim): test-only."""
    from cannoles_b200 import _capi
    hs = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-C", hs, "libb2_emu.so"], stdout=subprocess.DEVNULL)
    return _capi.bind_library(os.path.join(hs, "libb2_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib():
    from cannoles_b200 import _capi
    return _capi.load()
