"""Orthogonal Jacobi spaces (`FunctionSpace(N, 'J', alpha=, beta=)`, shenfun/jacobi/bases.py:150-260) against the
reference's own run (tests/golden/jacobi1d.npz): host-built operators on the CPU, the device contraction under
the gpu marker.  (alpha, beta) with alpha + beta = 0, alpha != 0 are left out: the reference's transforms
and its own `vandermonde` disagree there.)"""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
import shenfun_b200 as sf

CASES = [(0, 0), (1, 1), (0.5, 0.5), (2, 1), (0, 1.5)]
TOL = 1e-12


@pytest.mark.parametrize('ab', CASES)
@pytest.mark.parametrize('N', [16, 15])
def test_host_operators(ab, N):
    g = load_golden('jacobi1d')
    al, be = ab
    tag = f'{al}_{be}_{N}'
    S = sf.FunctionSpace(N, 'J', alpha=al, beta=be)
    x, w = S.points_and_weights()
    assert np.array_equal(x, g[f'x_{tag}']) and np.array_equal(w, g[f'wq_{tag}'])
    h = S.ortho_norm_sq()
    assert np.abs(h/g[f'h_{tag}']-1).max() < 1e-13
    V = S._vandermonde(N)
    sp = (V.T*w[None, :]) @ g[f'u_{tag}']
    assert rel_l2(sp, g[f'sp_{tag}']) < TOL
    assert rel_l2(sp/h, g[f'fw_{tag}']) < TOL
    assert rel_l2(V @ g[f'vh_{tag}'], g[f'bw_{tag}']) < TOL
    with pytest.raises(NotImplementedError):
        sf.FunctionSpace(N, 'J', alpha=al, beta=be, bc=(0, 0))


@pytest.mark.gpu
@pytest.mark.parametrize('ab', CASES)
@pytest.mark.parametrize('N', [16, 15])
def test_device_transforms(ab, N):
    g = load_golden('jacobi1d')
    al, be = ab
    tag = f'{al}_{be}_{N}'
    S = sf.FunctionSpace(N, 'J', alpha=al, beta=be)
    assert rel_l2(S.scalar_product(g[f'u_{tag}']).cpu().numpy(), g[f'sp_{tag}']) < TOL
    assert rel_l2(S.forward(g[f'u_{tag}']).cpu().numpy(), g[f'fw_{tag}']) < TOL
    assert rel_l2(S.backward(g[f'vh_{tag}']).cpu().numpy(), g[f'bw_{tag}']) < TOL
    Sp = S.get_dealiased(1.5)
    assert rel_l2(Sp.backward(g[f'vh_{tag}']).cpu().numpy(), g[f'bwpad_{tag}']) < TOL
    assert rel_l2(Sp.forward(g[f'w_{tag}']).cpu().numpy(), g[f'fwpad_{tag}']) < TOL


@pytest.mark.gpu
def test_jacobi_in_a_tensor_product_space():
    """Jacobi(1, 1) x Fourier^2: forward o backward is the identity, and alpha == beta operators take the
    parity-split contraction."""
    import torch
    from shenfun_b200 import ops
    J = sf.FunctionSpace(32, 'J', alpha=1, beta=1)
    assert ops._parity_halves(J._operators()[0]) is not None
    T = sf.TensorProductSpace(sf.comm, (J, sf.FunctionSpace(8, 'F', dtype='D'), sf.FunctionSpace(6, 'F', dtype='d')))
    u = sf.Array(T)
    u.v.normal_()
    uh = T.forward(u)
    ub = T.backward(uh)
    uh2 = T.forward(ub)
    assert (uh2.v-uh.v).abs().max().item() <= 1e-12*uh.v.abs().max().item()
    assert torch.isfinite(ub.v).all()
