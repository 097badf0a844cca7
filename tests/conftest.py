import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "gpu_next: GPU parity tests written when the round's GPU budget was spent -- not yet run on "
                                       "hardware, NOT part of -m gpu; run them with -m gpu_next and move them to `gpu` once green")


@pytest.fixture(scope="session")
def lib():
    from fastertransformer4codefuse_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no fallback path to test")
    from fastertransformer4codefuse_b200 import capi
    capi.check(capi.load().ftcf_device_check())
    return torch.device("cuda:0")
