"""pytest configuration: the `gpu` marker selects tests that need a real B200 (run by the driver with `-m gpu`)."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); everything else must pass on a CPU-only host")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "operators.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and requires a CUDA device; there is no CPU fallback")
    return torch.device("cuda:0")
