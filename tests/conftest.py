import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests.helpers import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product C-ABI library.  Fails loudly (no fallback) when it is missing or there is no device."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test collected without a CUDA device"
    import swimm_b200
    return swimm_b200
