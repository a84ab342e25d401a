import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_sm100():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:                                               # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a box without a B200 skips the gpu-marked tests instead of erroring 300 times
    (the product has no CPU path to run them on)."""
    if _have_sm100():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device of compute capability 10.x (B200); the product has no CPU path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as C
    C.build()
    return C
