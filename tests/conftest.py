import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are skipped (not failed) where they cannot run: no CUDA device, or libgom_b200.so not built.  On a GPU
    box with the library missing the product itself still raises GomError — only the test run degrades to skips."""
    import torch
    from gomavatar_b200 import _lib
    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device"
    elif not os.path.exists(_lib.LIB_PATH):
        reason = f"{_lib.LIB_PATH} has not been built (python -m gomavatar_b200.build)"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
