import os
import sys

import pytest

os.environ.setdefault("ZKB200_TEST_RNG", "1")      # the pinned-randomness hook of the library only exists in test processes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size reference comparisons")


def _has_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


@pytest.fixture(scope="session")
def zk():
    """The product library on cuda:0.  Fails loudly (no skip) when marked gpu tests run without the built library."""
    import blockmaze_b200 as zk
    zk.init(0)
    return zk


@pytest.fixture(scope="session")
def ref():
    from oracle import refapi
    if not refapi.available("kernels"):
        pytest.skip("oracle/_ref not built (needs /root/reference: run `make -C oracle`)")
    return refapi
