import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected on a box without a GPU."""
    return None


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    assert torch.cuda.is_available(), "this test is marked gpu and needs a CUDA device"
    return torch.device("cuda", 0)
