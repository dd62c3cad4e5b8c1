"""Legendre <-> Chebyshev conversion and DLT on the device (shenfun_b200.dlt) against the reference's
own vectors (tests/golden/dlt.npz) and, at bench-like lengths, against the host matrices."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from shenfun_b200 import dlt, ops

pytestmark = pytest.mark.gpu
TOL = 1e-12


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('N', [16, 33])
@pytest.mark.parametrize('axis', [0, 1])
def test_conversions_golden(N, axis):
    g = load_golden('dlt')
    x = g[f'x_{N}_{axis}']
    for key, fn in (('l2c', lambda a: dlt.leg2cheb(a, axis=axis)), ('l2cT', lambda a: dlt.leg2cheb(a, axis=axis, transpose=True)),
                    ('c2l', lambda a: dlt.cheb2leg(a, axis=axis))):
        ref = g[f'{key}_{N}_{axis}']
        got = fn(dev(x)).cpu().numpy()
        assert np.abs(got-ref).max() <= TOL*np.abs(ref).max(), key
        host = fn(x.copy())                                  # numpy in, numpy out (staged through the device)
        assert isinstance(host, np.ndarray) and np.abs(host-ref).max() <= TOL*np.abs(ref).max()
    xr = g[f'xr_{N}']
    assert np.abs(dlt.leg2cheb(dev(xr)).cpu().numpy()-g[f'l2c_r_{N}']).max() <= TOL*np.abs(g[f'l2c_r_{N}']).max()
    assert np.abs(dlt.Cheb2Leg(dev(xr))(dev(xr)).cpu().numpy()-g[f'c2l_r_{N}']).max() <= TOL*np.abs(g[f'c2l_r_{N}']).max()


@pytest.mark.parametrize('N', [256, 512, 1000])
def test_long_lines_against_the_host_matrix_and_round_trip(N):
    """Checkerboard tensor-core path with the zero k-tiles skipped (upper triangular halves; the transposed
    operator is lower triangular)."""
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2, N, 37))+1j*rng.standard_normal((2, N, 37))
    M, L = dlt._l2c_matrix(N), dlt._c2l_matrix(N)
    A = ops.as_device_tensor(M)
    assert ops._parity_halves(A)[1] == (2 | 4)
    assert ops._parity_halves(ops.as_device_tensor(np.ascontiguousarray(M.T)))[1] == (2 | 8)
    xd = dev(x)
    c = dlt.FMMLeg2Cheb(xd, axis=1)(xd)
    ref = np.einsum('ij,ojc->oic', M, x)
    assert np.abs(c.cpu().numpy()-ref).max() <= TOL*np.abs(ref).max()
    ct = dlt.leg2cheb(xd, axis=1, transpose=True)
    reft = np.einsum('ji,ojc->oic', M, x)
    assert np.abs(ct.cpu().numpy()-reft).max() <= TOL*np.abs(reft).max()
    back = dlt.FMMCheb2Leg(c, axis=1)(c)
    # the conversion pair is ill conditioned like N^(1/2) in the max norm; 1e-12 relative holds to N = 1000
    assert np.abs(back.cpu().numpy()-x).max() <= 1e-12*np.abs(x).max()*max(1.0, N/256)


@pytest.mark.parametrize('N', [16, 33])
def test_dlt_golden(N):
    g = load_golden('dlt')
    uh = g[f'dlt_uh_{N}']
    ub = dlt.DLT(dev(uh), kind='backward')()
    assert np.abs(ub.cpu().numpy()-g[f'dlt_bwd_{N}']).max() <= TOL*np.abs(g[f'dlt_bwd_{N}']).max()
    sp = dlt.DLT(ub, kind='scalar product')()
    assert np.abs(sp.cpu().numpy()-g[f'dlt_sp_{N}']).max() <= TOL*np.abs(g[f'dlt_sp_{N}']).max()
    fw = dlt.DLT(ub, kind='forward')()
    assert np.abs(fw.cpu().numpy()-uh).max() <= TOL*np.abs(uh).max()
