import os
import subprocess
import sys

import pytest

# In-process logical stripes (tests/test_gpu_stripes.py) run several engines' streams with
# cross-stream spin-wait kernels on ONE GPU: give every stream its own hardware queue so a waiting
# kernel can never sit in front of the kernel it waits for (must be set before CUDA initialises;
# irrelevant for the real one-process-per-GPU layout).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer-running CPU test")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the oracle port exist (nvcc cross-compiles on CPU)."""
    import __graft_entry__ as ge
    ge.build_cuda()
    ge.build_oracle()
    yield
