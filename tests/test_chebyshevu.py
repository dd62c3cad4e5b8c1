"""Orthogonal Chebyshev spaces of the second kind (`FunctionSpace(N, 'U')`, shenfun/chebyshevu/bases.py) against
the reference's own run (tests/golden/chebyshevu1d.npz); padded transforms are checked by a round trip only --
the reference's own padded backward of this family raises (bases.py:252)."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
import shenfun_b200 as sf

TOL = 1e-12


@pytest.mark.parametrize('quad', ['GU', 'GC'])
@pytest.mark.parametrize('N', [16, 15])
def test_host_operators(quad, N):
    g = load_golden('chebyshevu1d')
    tag = f'{quad}_{N}'
    S = sf.FunctionSpace(N, 'U', quad=quad)
    x, w = S.points_and_weights()
    assert np.array_equal(x, g[f'x_{tag}']) and np.array_equal(w, g[f'wq_{tag}'])
    V = S._vandermonde(N)
    sp = (V.T*w[None, :]) @ g[f'u_{tag}']
    assert rel_l2(sp, g[f'sp_{tag}']) < TOL
    assert rel_l2(sp/S.ortho_norm_sq(), g[f'fw_{tag}']) < TOL
    assert rel_l2(V @ g[f'vh_{tag}'], g[f'bw_{tag}']) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('quad', ['GU', 'GC'])
@pytest.mark.parametrize('N', [16, 15])
def test_device_transforms(quad, N):
    g = load_golden('chebyshevu1d')
    tag = f'{quad}_{N}'
    S = sf.FunctionSpace(N, 'U', quad=quad)
    assert rel_l2(S.scalar_product(g[f'u_{tag}']).cpu().numpy(), g[f'sp_{tag}']) < TOL
    assert rel_l2(S.forward(g[f'u_{tag}']).cpu().numpy(), g[f'fw_{tag}']) < TOL
    assert rel_l2(S.backward(g[f'vh_{tag}']).cpu().numpy(), g[f'bw_{tag}']) < TOL
    if quad == 'GU':
        Sp = S.get_dealiased(1.5)          # exact quadrature: forward o backward = identity on the N modes
        vh = g[f'vh_{tag}']
        assert rel_l2(Sp.forward(Sp.backward(vh)).cpu().numpy(), vh) < TOL
