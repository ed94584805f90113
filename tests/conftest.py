import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _gpu_unavailable_reason():
    """None when the `gpu` tests must run: a CUDA device is present.  (A missing libstyle_b200.so on
    a GPU machine is NOT a reason to skip: the tests then fail loudly, as the product path does.)"""
    try:
        import torch
        if not torch.cuda.is_available():
            return 'no CUDA device (the engine has no CPU path)'
    except Exception as e:                                   # pragma: no cover
        return 'torch unavailable: %r' % (e,)
    return None


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped, not failed, on a machine without a GPU, so that a plain
    `pytest tests` is green on CPU-only CI."""
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
