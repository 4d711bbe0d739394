import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # An explicit `-m gpu` run fails loudly (not skip) when there is no CUDA device: no silent fallback.  A plain `pytest` on a
    # host without CUDA deselects the gpu-marked tests instead of turning the run red (`-m "not gpu"` is what CI runs there).
    if config.getoption("-m"):
        return
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    keep, drop = [], []
    for it in items:
        (drop if it.get_closest_marker("gpu") else keep).append(it)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep
