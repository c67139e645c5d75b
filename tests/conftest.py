import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        # A kernel that never returns (a lost arrival in one of the samplers' barrier protocols) must fail the run, not hang the
        # box: the thread method dumps the stacks and ends the process even while the main thread sits in cudaDeviceSynchronize.
        if has_timeout and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(180 if "gpu" in item.keywords else 900, method="thread"))
