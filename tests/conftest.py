import os
import sys

import pytest

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # as the package does on import; here before torch probes the device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests must fail loudly (not skip) when the device or the CUDA library is missing;
    # plain runs without a marker expression on a CPU box skip them.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
