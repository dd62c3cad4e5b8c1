"""Legendre bases on the Gauss-Lobatto points (quad='GL'): nodes, weights, mass matrices and the host-built
operator matrices against the reference's own run (tests/golden/poly1d_LGL.npz, legendre/bases.py:128-129,
157-164).  The device contraction of the same operators is checked in tests/test_gpu_poly.py."""
import numpy as np
import pytest

from conftest import load_golden
import shenfun_b200 as sf
from shenfun_b200.legendre import leglobatto

BCS = {'dirichlet': (0, 0), 'biharmonic': (0, 0, 0, 0), 'neumann': {'left': {'N': 0}, 'right': {'N': 0}}, 'ortho': None}


def dense(M):
    d = M.diags() if hasattr(M, 'diags') else M
    return np.asarray(d.toarray() if hasattr(d, 'toarray') else d)


@pytest.mark.parametrize('N', [16, 15])
def test_nodes_and_weights(N):
    g = load_golden('poly1d_LGL')
    x, w = leglobatto(N)
    assert np.abs(x-g[f'x_ortho_{N}']).max() < 1e-15
    assert np.abs(w-g[f'wq_ortho_{N}']).max() < 1e-15


@pytest.mark.parametrize('name', list(BCS))
@pytest.mark.parametrize('N', [16, 15])
def test_operators_and_mass(name, N):
    g = load_golden('poly1d_LGL')
    bc = BCS[name]
    S = sf.FunctionSpace(N, 'L', bc=bc, quad='GL') if bc is not None else sf.FunctionSpace(N, 'L', quad='GL')
    tag = f'{name}_{N}'
    d = S.dim()
    K = dense(S.stencil_matrix()) if bc is not None else np.eye(N)
    for sp_, ukey, fkey in ((S, 'u', 'fw'), (S.get_dealiased(1.5), 'w', 'fwpad')):
        Np = sp_.physical_len()
        x, w = sp_.points_and_weights(Np)
        V = sp_._vandermonde(Np)
        sp = (K @ (V.T*w[None, :])) @ g[f'{ukey}_{tag}']
        fw = np.linalg.solve(dense(sp_.mass_matrix()), sp[:d])
        assert np.abs(fw-g[f'{fkey}_{tag}'][:d]).max() < 1e-12
    if f'M00_{tag}' in g.files:
        assert np.abs(dense(S.mass_matrix())-g[f'M00_{tag}']).max() < 1e-14
