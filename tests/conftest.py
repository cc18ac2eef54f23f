import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU path.
    # (Without -m, gpu tests are skipped on GPU-less machines so that a bare `pytest` stays usable.)
    if config.getoption('-m'):
        return
    if not _have_gpu():
        skip = pytest.mark.skip(reason='no CUDA device')
        for item in items:
            if 'gpu' in item.keywords:
                item.add_marker(skip)
