"""`shenfun_b200.fftw`: the mpi4py_fft.fftw plan surface on the device, against scipy.fft
(pocketfft, the same un-normalised FFTW conventions).  Tolerance 1e-12 relative L2."""
import numpy as np
import pytest
import scipy.fft as sfft
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rand(shape, cplx, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(shape)
    return a+1j*rng.standard_normal(shape) if cplx else a


@pytest.mark.parametrize('shape,axes', [((16, 12, 10), (2,)), ((16, 12, 10), (1,)), ((16, 12, 10), (0,)),
                                        ((16, 12, 10), (0, 1, 2)), ((9, 64), (0, 1)), ((128,), (0,))])
def test_c2c_plans(shape, axes):
    from shenfun_b200 import fftw
    a = _rand(shape, True, 1)
    U = fftw.aligned(shape, dtype='D')
    fwd = fftw.fftn(U, axes=axes, threads=2, flags=(fftw.FFTW_MEASURE, fftw.FFTW_DESTROY_INPUT))
    V = fwd.output_array
    bck = fftw.ifftn(V, axes=axes, output_array=U)
    U.copy_(torch.from_numpy(a).cuda())
    out = fwd()
    assert out is fwd.output_array
    assert rel_l2(out.cpu().numpy(), sfft.fftn(a, axes=axes)) < TOL
    assert abs(fwd.get_normalization()-1.0/np.prod([shape[i] for i in axes])) < 1e-18
    back = bck()
    assert rel_l2(back.cpu().numpy()*fwd.get_normalization(), a) < TOL
    # explicit arrays + normalize
    res = torch.empty_like(V)
    fwd(torch.from_numpy(a).cuda(), res, normalize=True)
    assert rel_l2(res.cpu().numpy(), sfft.fftn(a, axes=axes, norm='forward')) < TOL


@pytest.mark.parametrize('shape,axes', [((16, 12, 10), (2,)), ((16, 12, 10), (0, 1, 2)), ((12, 9, 16), (1, 2)),
                                        ((8, 16), (1, 0)), ((64,), (0,))])
def test_r2c_plans(shape, axes):
    from shenfun_b200 import fftw
    a = _rand(shape, False, 2)
    U = fftw.aligned(shape, dtype=float)
    s = [shape[i] for i in axes]
    fwd = fftw.rfftn(U, s=s, axes=axes)
    V = fwd.output_array
    bck = fftw.irfftn(V, s=s, axes=axes, output_array=U)
    U.copy_(torch.from_numpy(a).cuda())
    ref = sfft.rfftn(a, axes=axes)
    assert tuple(V.shape) == ref.shape
    assert rel_l2(fwd().cpu().numpy(), ref) < TOL
    assert rel_l2(bck().cpu().numpy()*fwd.get_normalization(), a) < TOL


@pytest.mark.parametrize('cplx', (False, True))
@pytest.mark.parametrize('axis', (0, 1, 2))
def test_dct_plans(cplx, axis):
    from shenfun_b200 import fftw
    shape = (12, 16, 10)
    a = _rand(shape, cplx, 3)
    U = fftw.aligned(shape, dtype='D' if cplx else 'd')
    fwd = fftw.dctn(U, axes=(axis,), type=2)
    bck = fftw.dctn(fwd.output_array, axes=(axis,), type=3, output_array=U)
    U.copy_(torch.from_numpy(a).cuda())

    def dct(x, t):
        if np.iscomplexobj(x):
            return sfft.dct(x.real, type=t, axis=axis)+1j*sfft.dct(x.imag, type=t, axis=axis)
        return sfft.dct(x, type=t, axis=axis)
    assert rel_l2(fwd().cpu().numpy(), dct(a, 2)) < TOL
    assert rel_l2(bck().cpu().numpy(), dct(dct(a, 2), 3)) < TOL
    assert abs(fwd.get_normalization()-1.0/(2*shape[axis])) < 1e-18
    with pytest.raises(NotImplementedError):
        fftw.dctn(U, axes=(axis,), type=1)
