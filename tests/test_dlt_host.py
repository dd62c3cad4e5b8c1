"""Connection matrices of shenfun_b200.dlt (host side) against vectors produced by the reference's
legendre/dlt.py (tests/golden/make_golden.py legendre_chebyshev_conversion) and, when oracle/_ref holds
them, against the reference's compiled kernels (optimization/cython/transforms.pyx)."""
import numpy as np
import pytest

from conftest import load_golden
from shenfun_b200 import dlt

TOL = 1e-12


@pytest.mark.parametrize('N', [16, 33])
@pytest.mark.parametrize('axis', [0, 1])
def test_connection_matrices_golden(N, axis):
    g = load_golden('dlt')
    x = np.moveaxis(g[f'x_{N}_{axis}'], axis, 0)
    for key, M in (('l2c', dlt._l2c_matrix(N)), ('l2cT', dlt._l2c_matrix(N).T), ('c2l', dlt._c2l_matrix(N))):
        ref = np.moveaxis(g[f'{key}_{N}_{axis}'], axis, 0)
        got = np.tensordot(M, x, axes=(1, 0))
        assert np.abs(got-ref).max() <= TOL*np.abs(ref).max(), key


def test_matrices_are_checkerboard_upper_triangular_and_inverse():
    N = 64
    M, L = dlt._l2c_matrix(N), dlt._c2l_matrix(N)
    for A in (M, L):
        assert not np.tril(A, -1).any()
        assert not A[0::2, 1::2].any() and not A[1::2, 0::2].any()
    assert np.abs(L @ M-np.eye(N)).max() < 1e-12


def test_against_reference_compiled_kernels():
    from oracle import refcy
    if not refcy.available():
        pytest.skip('oracle/_ref/refcy not built')
    tr = refcy.load('transforms')
    rng = np.random.default_rng(3)
    N = 40
    x = rng.standard_normal((N, 3))+1j*rng.standard_normal((N, 3))
    ref = tr.leg2cheb(x.copy(), np.zeros_like(x), 0, 0)
    assert np.abs(dlt._l2c_matrix(N) @ x-ref).max() <= TOL*np.abs(ref).max()
    ref = tr.cheb2leg(x.copy(), np.zeros_like(x), 0)
    assert np.abs(dlt._c2l_matrix(N) @ x-ref).max() <= TOL*np.abs(ref).max()
