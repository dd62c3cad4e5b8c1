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


BCS = {
    'ortho': None,
    'dirichlet': (0, 0),
    'neumann': {'left': {'N': 0}, 'right': {'N': 0}},
    'biharmonic': (0, 0, 0, 0),
    'dn': {'left': {'D': 0}, 'right': {'N': 0}},
    'inhom': {'left': {'D': 1.5}, 'right': {'D': -0.5}},
}


def oracle_spec(base):
    """What oracle/polybases.py needs, taken from the host-side basis object (whose stencil and
    mass matrix are themselves pinned against the reference in test_oracle_poly.py)."""
    spec = dict(family='C' if base.family() == 'chebyshev' else 'L', N=base.N, n_phys=base.physical_len(),
                dim=base.dim(), stencil=base.stencil(), df=float(base.domain_factor()),
                mass=base.mass_matrix().diags())
    if spec['family'] == 'L':
        spec['x'], spec['w'] = base.points_and_weights(base.physical_len())
    return spec
