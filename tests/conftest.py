import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is test infrastructure; (re)build it before any test touches it."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    yield


def cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
