"""Parity-split tensor-core contraction (`shb_gemm_axis_parity`, csrc/gemm.cu) against numpy: operators
with the even/odd structure of the symmetric families, both directions, odd and even sizes, real and
complex data; operators without the structure must take the dense kernel."""
import numpy as np
import pytest
import torch

from shenfun_b200 import ops
import shenfun_b200 as sf

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _structured(rng, M, K, forward):
    a = rng.standard_normal((M, K))
    if forward:      # A[i, K-1-j] = (-1)^i A[i, j]
        sgn = (-1.0)**np.arange(M)[:, None]
        a[:, K//2:] = (sgn*a[:, ::-1])[:, K//2:]
    else:            # A[M-1-i, k] = (-1)^k A[i, k]
        sgn = (-1.0)**np.arange(K)[None, :]
        a[M//2:] = (sgn*a[::-1])[M//2:]
    return a


@pytest.mark.parametrize('forward', [True, False])
@pytest.mark.parametrize('M,K', [(256, 256), (254, 256), (256, 254), (130, 96), (64, 200), (96, 66), (37, 64), (64, 37)])
@pytest.mark.parametrize('outer,inner,cplx', [(1, 129, True), (3, 40, False), (2, 1, True), (1, 1000, True)])
def test_parity_contraction(forward, M, K, outer, inner, cplx):
    if (forward and K % 2) or (not forward and M % 2):
        pytest.skip('structure needs an even mirrored extent')
    rng = np.random.default_rng(M*1000+K+inner)
    a = _structured(rng, M, K, forward)
    A = ops.as_device_tensor(a)
    halves = ops._parity_halves(A)
    assert halves is not None and halves[1] == int(forward)
    u = rng.standard_normal((outer, K, inner))
    if cplx:
        u = u+1j*rng.standard_normal((outer, K, inner))
    ref = np.einsum('ij,ojc->oic', a, u)
    ud = torch.from_numpy(u).cuda()
    out = torch.full((outer, M, inner), 7.0, dtype=ud.dtype, device='cuda')
    n0 = ops.LAUNCHES
    ops.gemm_axis(A, ud, out, outer, inner)
    assert ops.LAUNCHES == n0+1
    err = np.abs(out.cpu().numpy()-ref).max()/np.abs(ref).max()
    assert err < TOL, err


def test_unstructured_operator_takes_the_dense_kernel():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((64, 64))
    A = ops.as_device_tensor(a)
    assert ops._parity_halves(A) is None
    u = torch.randn(2, 64, 33, dtype=torch.complex128, device='cuda')
    out = torch.empty_like(u)
    ops.gemm_axis(A, u, out, 2, 33)
    ref = np.einsum('ij,ojc->oic', a, u.cpu().numpy())
    assert np.abs(out.cpu().numpy()-ref).max()/np.abs(ref).max() < TOL


@pytest.mark.parametrize('bc', [(0, 0), (0, 0, 0, 0), None])
def test_legendre_operators_have_the_structure(bc):
    """The Legendre operator matrices of homogeneous spaces are parity-structured (so the transforms run
    at half the dense cost); an inhomogeneous space keeps its boundary columns and the dense kernel."""
    S = sf.FunctionSpace(64, 'L', bc=bc) if bc is not None else sf.FunctionSpace(64, 'L')
    F, B = S._operators()
    assert ops._parity_halves(F) is not None and ops._parity_halves(F)[1] == 1
    assert ops._parity_halves(B) is not None and ops._parity_halves(B)[1] == 0
    Si = sf.FunctionSpace(64, 'L', bc=(1.0, -2.0))
    assert ops._parity_halves(Si._operators()[1]) is None
