"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic,
C-ABI symbol export); `-m gpu` are the parity tests proper and need a B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def exp868():
    return dict(np.load(os.path.join(GOLDEN, 'exp868.npz')))


@pytest.fixture(scope='session')
def dirichlet_golden():
    return dict(np.load(os.path.join(GOLDEN, 'dirichlet_fit.npz')))
