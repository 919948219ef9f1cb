import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def backend():
    """The CUDA backend through its C ABI. Fails loudly (no fallback) if the library or the GPU is missing."""
    import torch

    from spla_b200.backend import Backend

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return Backend(0)
