import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def fixture_sd():
    from oracle import fixtures as FX
    return FX.make_state_dict(0)


@pytest.fixture(scope='session')
def golden_small():
    return dict(np.load(os.path.join(GOLDEN, 'small.npz')))


@pytest.fixture(scope='session')
def golden_full():
    return dict(np.load(os.path.join(GOLDEN, 'full.npz')))
