import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_path():
    from posegen_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def engine(lib_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from posegen_b200.engine import Engine
    return Engine()
