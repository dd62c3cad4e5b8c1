"""Dense (matrix) transforms of the orthogonal-polynomial families under the reference's names:
`scalar_product(input, output, x, w, axis, a)` and `evaluate_expansion_all(input, output, x,
axis, a)` (imported by path at shenfun/spectralbase.py:485-486, 552-553; kernels
optimization/cython/transforms.pyx:175-323).

`a` holds the three-term recurrence of the family, x P_n = a[0][n+1]... in the reference's
convention  P_{n+1} = ((x - ann_n) P_n - anp_n P_{n-1}) / anm_n  with rows (anm, ann, anp) or
(anm, anp) when ann = 0.  The reference re-runs that recurrence for every line; here it runs
once on the host (extended precision) into the operator matrix and the contraction is one
fp64 tensor-core GEMM (shb_gemm_axis).
"""
import numpy as np
import torch

from ... import ops
from ._args import Staged, axis_view, host

_CACHE = {}


def _vandermonde(x, a, M):
    """V[k][j] = P_k(x_j), k < M."""
    x = np.asarray(host(x), dtype=np.longdouble)
    a = np.asarray(host(a), dtype=np.longdouble)
    if a.shape[0] == 3:
        anm, ann, anp = a[0], a[1], a[2]
    else:
        anm, anp = a[0], a[1]
        ann = np.zeros_like(anm)
    V = np.zeros((M, x.shape[0]), dtype=np.longdouble)
    V[0] = 1
    if M > 1:
        V[1] = (x-ann[0])/anm[0]
    for n in range(1, M-1):
        V[n+1] = ((x-ann[n])*V[n]-anp[n]*V[n-1])/anm[n]
    return V


def _operator(kind, x, w, a, M):
    xa, aa = np.ascontiguousarray(host(x)), np.ascontiguousarray(host(a))
    wa = None if w is None else np.ascontiguousarray(host(w))
    key = (kind, M, xa.tobytes(), aa.tobytes(), None if wa is None else wa.tobytes())
    op = _CACHE.get(key)
    if op is None:
        V = _vandermonde(xa, aa, M)
        if kind == 'scalar_product':
            A = V*np.asarray(wa, dtype=np.longdouble)[None, :]       # [M][N]
        else:
            A = V.T                                                  # [N][M]
        op = ops.as_device_tensor(np.ascontiguousarray(A.astype(np.float64)))
        if len(_CACHE) > 16:
            _CACHE.clear()
        _CACHE[key] = op
    return op


def scalar_product(input_array, output_array, x, w, axis, a):
    """output_k = sum_j w_j input_j P_k(x_j) along `axis`, k < output.shape[axis]."""
    st = Staged()
    out = st(output_array, out=True)
    inp = st(input_array, dtype=out.dtype)
    outer, M, inner = axis_view(out.shape, axis)
    A = _operator('scalar_product', x, w, a, M)
    assert inp.shape[axis] == A.shape[1]
    ops.gemm_axis(A, inp, out, outer, inner)
    st.finish()
    return output_array


def evaluate_expansion_all(input_array, output_array, x, axis, a):
    """output_j = sum_k input_k P_k(x_j) along `axis`."""
    st = Staged()
    out = st(output_array, out=True)
    inp = st(input_array, dtype=out.dtype)
    outer, N, inner = axis_view(out.shape, axis)
    M = inp.shape[axis]
    A = _operator('evaluate_expansion_all', x, None, a, M)
    assert N == A.shape[0]
    ops.gemm_axis(A, inp, out, outer, inner)
    st.finish()
    return output_array
