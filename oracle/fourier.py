"""Oracle: Fourier-axis transforms with shenfun's pad/truncate/normalise rules.

Restates (numpy + scipy.fft/pocketfft in fp64):
  shenfun/fourier/bases.py:131-149  FourierBase.forward  (FFT, truncate, *1/N_phys)
  shenfun/fourier/bases.py:404-435  R2C backward / _truncation_forward / _padding_backward
  shenfun/fourier/bases.py:530-598  C2C backward / truncation / padding
FFTW conventions (mpi4py_fft.fftw, third-party, absent): forward and backward
transforms are un-normalised; get_normalization() = 1/n.
"""
import numpy as np
import scipy.fft as sfft

WORKERS = 1


def _sl(ndim, axis, s):
    out = [slice(None)]*ndim
    out[axis] = s
    return tuple(out)


def c2c_forward(u, axis, N, n_fft=None, workers=None):
    """fourier/bases.py:131-149 with C2C._truncation_forward :570-577."""
    n_fft = u.shape[axis] if n_fft is None else n_fft
    assert u.shape[axis] == n_fft
    X = sfft.fft(u, axis=axis, workers=workers or WORKERS)
    if N != n_fft:
        sh = list(u.shape)
        sh[axis] = N
        out = np.zeros(sh, dtype=complex)
        out[_sl(u.ndim, axis, slice(0, N//2+1))] = X[_sl(u.ndim, axis, slice(0, N//2+1))]
        out[_sl(u.ndim, axis, slice(-(N//2), None))] += X[_sl(u.ndim, axis, slice(-(N//2), None))]
        X = out
    return X*(1./n_fft)


def c2c_backward(uh, axis, n_fft=None, dealias_direct=False, workers=None):
    """fourier/bases.py:530-533 with C2C._padding_backward :579-598."""
    N = uh.shape[axis]
    n_fft = N if n_fft is None else n_fft
    if n_fft != N:
        sh = list(uh.shape)
        sh[axis] = n_fft
        p = np.zeros(sh, dtype=complex)
        p[_sl(uh.ndim, axis, slice(0, N//2+1))] = uh[_sl(uh.ndim, axis, slice(0, N//2+1))]
        p[_sl(uh.ndim, axis, slice(-(N//2), None))] = uh[_sl(uh.ndim, axis, slice(-(N//2), None))]
        if N % 2 == 0:
            p[_sl(uh.ndim, axis, N//2)] *= 0.5
            p[_sl(uh.ndim, axis, -(N//2))] *= 0.5
    else:
        p = np.array(uh, dtype=complex)
        if dealias_direct:
            p[_sl(uh.ndim, axis, slice(N//3, -(N//3)+1))] = 0
    return sfft.ifft(p, axis=axis, norm='forward', workers=workers or WORKERS)


def r2c_forward(u, axis, N, n_fft=None, workers=None):
    """fourier/bases.py:131-149 with R2C._truncation_forward :409-418."""
    n_fft = u.shape[axis] if n_fft is None else n_fft
    assert u.shape[axis] == n_fft
    X = sfft.rfft(u, axis=axis, workers=workers or WORKERS)
    if N != n_fft:
        ns = N//2+1
        X = np.array(X[_sl(u.ndim, axis, slice(0, ns))])
        if N % 2 == 0:
            s1 = _sl(u.ndim, axis, ns-1)
            X[s1] = X[s1].real*2
    return X*(1./n_fft)


def c2r_backward(uh, axis, N, n_fft=None, dealias_direct=False, workers=None):
    """fourier/bases.py:404-407 with R2C._padding_backward :420-435."""
    ns = uh.shape[axis]
    assert ns == N//2+1
    n_fft = N if n_fft is None else n_fft
    if n_fft != N:
        sh = list(uh.shape)
        sh[axis] = n_fft//2+1
        p = np.zeros(sh, dtype=complex)
        p[_sl(uh.ndim, axis, slice(0, ns))] = uh
        if N % 2 == 0:
            sm = _sl(uh.ndim, axis, ns-1)
            p[sm] = p[sm].real*0.5
    else:
        p = np.array(uh, dtype=complex)
        if dealias_direct:
            p[_sl(uh.ndim, axis, slice(N//3, None))] = 0
    return sfft.irfft(p, n=n_fft, axis=axis, norm='forward', workers=workers or WORKERS)
