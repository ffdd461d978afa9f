import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_backend():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from asva_b200 import _lib, ops

    lib = _lib.load()  # raises if the .so is missing: no fallback
    _lib.check(lib.asva_device_check(), "asva_device_check")
    return ops.CudaBackend()
