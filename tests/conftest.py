import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs /root/reference and the oracle/_ref build (build container only)')


@pytest.fixture(scope='session')
def golden_dir():
    return ROOT / 'tests' / 'golden'


def pytest_collection_modifyitems(config, items):
    """GPU tests need a CUDA device: without one they are skipped rather than failed (the driver selects with -m anyway)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
