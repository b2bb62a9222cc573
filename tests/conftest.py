import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def addition_family():
    from vael_b200.mnist_task import compile_addition_family
    return compile_addition_family(2, 10)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def _seeded_default_generators():
    """Every test starts from the same default CPU / CUDA generator state: inputs drawn without an explicit generator are
    the same on every run, so a pass or a failure is reproducible."""
    import torch
    torch.manual_seed(20240)
    yield
