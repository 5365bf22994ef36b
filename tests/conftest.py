import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.build()
    return oracle_lib


def pytest_collection_modifyitems(config, items):
    """`-m gpu` on a machine without a CUDA device: skip instead of erroring inside every test."""
    if not any("gpu" in it.keywords for it in items):
        return
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product path has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
