import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a)
    b = np.asarray(b)
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a-b).ravel())/(nb if nb > 0 else 1.0)


@pytest.fixture(scope='session')
def lib():
    from shenfun_b200 import _lib
    return _lib.load()
