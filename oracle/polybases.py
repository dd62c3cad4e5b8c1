"""Oracle: Chebyshev / Legendre composite-basis transforms along one axis.

Restates (numpy + scipy.fft):
  shenfun/chebyshev/bases.py:315-353   Orthogonal._evaluate_expansion_all / _evaluate_scalar_product (GC)
  shenfun/spectralbase.py:200-313      scalar_product / forward / backward skeleton
  shenfun/spectralbase.py:1363-1390    polynomial truncation / padding (first dim() modes)
  shenfun/spectralbase.py:1462-1474    composite scalar product  out = K w
  shenfun/spectralbase.py:1616-1634    to_ortho  w = K^T u_hat
  shenfun/spectralbase.py:560-577      apply_inverse_mass (banded solve along the axis)
  shenfun/spectralbase.py:473-483,539-548  'vandermonde' Legendre evaluation / scalar product
A `spec` dict carries what the reference derives from the basis object:
  family ('C' or 'L'), N, n_phys, dim, stencil {offset: values[dim]}, df (domain factor),
  mass (dense dim x dim mass matrix of the composite basis), x / w (Legendre nodes, weights).
"""
import numpy as np
import scipy.fft as sfft
import scipy.linalg as sla


def stencil_dense(spec, ncols=None):
    dim = spec['dim']
    ncols = spec['N'] if ncols is None else ncols
    K = np.zeros((dim, ncols))
    for off, v in spec['stencil'].items():
        v = np.broadcast_to(v, (dim,))
        for k in range(dim):
            if k+off < ncols:
                K[k, k+off] = v[k]
    return K


def _apply(M, u, axis):
    """Dense M along `axis`."""
    return np.moveaxis(np.tensordot(M, u, axes=([1], [axis])), 0, axis)


def legendre_vandermonde(x, N):
    V = np.zeros((len(x), N))
    V[:, 0] = 1
    if N > 1:
        V[:, 1] = x
    for n in range(2, N):
        V[:, n] = ((2*n-1)*x*V[:, n-1]-(n-1)*V[:, n-2])/n
    return V


def scalar_product_axis(u, axis, spec):
    N, Np, dim = spec['N'], spec['n_phys'], spec['dim']
    assert u.shape[axis] == Np
    if spec['family'] == 'C':
        # chebyshev/bases.py:345-347: DCT-II * pi/(2 df Np)   (complex: real and imaginary separately)
        if np.iscomplexobj(u):
            w = sfft.dct(u.real, type=2, axis=axis)+1j*sfft.dct(u.imag, type=2, axis=axis)
        else:
            w = sfft.dct(u, type=2, axis=axis)
        w = w*(np.pi/(2*spec['df']*Np))
    else:
        V = legendre_vandermonde(spec['x'], Np)           # ortho modes up to Np
        w = _apply((V.T*spec['w'][None, :])/spec['df'], u, axis)
    K = stencil_dense(spec, Np)                            # spectralbase.py:1471 stencil_matrix(shape(False))
    s = _apply(K, w, axis)
    out_shape = list(u.shape)
    out_shape[axis] = N
    out = np.zeros(out_shape, dtype=s.dtype)
    sl = [slice(None)]*u.ndim
    sl[axis] = slice(0, dim)
    out[tuple(sl)] = s
    return out


def forward_axis(u, axis, spec, scalar_product=False):
    out = scalar_product_axis(u, axis, spec)
    if scalar_product:
        return out
    dim = spec['dim']
    sl = [slice(None)]*u.ndim
    sl[axis] = slice(0, dim)
    b = np.moveaxis(out[tuple(sl)], axis, 0)
    x = sla.solve(spec['mass'], b.reshape(dim, -1)).reshape(b.shape)
    out[tuple(sl)] = np.moveaxis(x, 0, axis)
    return out


def backward_axis(uh, axis, spec):
    N, Np, dim = spec['N'], spec['n_phys'], spec['dim']
    assert uh.shape[axis] == N
    sl = [slice(None)]*uh.ndim
    sl[axis] = slice(0, dim)
    K = stencil_dense(spec, Np if Np >= N else N)
    c = _apply(K.T, uh[tuple(sl)], axis)                   # ortho coefficients, zero padded to Np
    if spec['family'] == 'C':
        # chebyshev/bases.py:326-329: 0.5*DCT-III + c_0/2
        if np.iscomplexobj(c):
            y = sfft.dct(c.real, type=3, axis=axis)+1j*sfft.dct(c.imag, type=3, axis=axis)
        else:
            y = sfft.dct(c, type=3, axis=axis)
        s0 = [slice(None)]*uh.ndim
        s0[axis] = slice(0, 1)
        return 0.5*y+c[tuple(s0)]/2
    V = legendre_vandermonde(spec['x'], K.shape[1])
    return _apply(V, c, axis)
