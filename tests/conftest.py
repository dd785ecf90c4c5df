import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_available():
    try:
        import torch

        if not torch.cuda.is_available():
            return False
        from physicl_b200 import _capi

        _capi.Context(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest tests`` on a box without a usable B200 skips the GPU tests instead of failing them."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libphysicl_b200.so")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden
