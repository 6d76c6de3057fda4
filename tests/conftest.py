import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: more than a few seconds on the CPU")


def _have_cuda_device():
    if not os.path.exists("/dev/nvidia0") and not os.path.exists("/dev/nvidiactl"):
        return False
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a machine without a B200 skips the gpu-marked tests instead of failing in them."""
    if _have_cuda_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product path has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def libepic_built():
    """The product library, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from epic_b200 import libepic
    if not os.path.exists(libepic.LIB_PATH):
        libepic.build()
    return libepic.load()
