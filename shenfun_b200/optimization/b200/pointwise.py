"""Pointwise kernels under the reference's names (optimization/cython/cross.pyx:25-40,
applymask.pyx:15-169, outer.pyx)."""
import numpy as np
import torch

from ... import ops
from ._args import Staged


def cross3D(c, a, b):
    """c = a x b on (3,) + shape arrays."""
    st = Staged()
    ct = st(c, out=True)
    at = st(a, dtype=ct.dtype)
    bt = st(b, dtype=ct.dtype)
    ops.cross3d(at, bt, ct)
    st.finish()
    return c


def cross2D(c, a, b):
    """c = a0*b1 - a1*b0 on (2,) + shape arrays (scalar result)."""
    st = Staged()
    ct = st(c, out=True)
    at = st(a, dtype=ct.dtype)
    bt = st(b, dtype=ct.dtype)
    torch.mul(at[0], bt[1], out=ct)
    ct.addcmul_(at[1], bt[0], value=-1)
    st.finish()
    return c


def _outer(a, b, c, symmetric, dim):
    st = Staged()
    ct = st(c, out=True)
    at = st(a, dtype=ct.dtype)
    bt = at if symmetric else st(b, dtype=ct.dtype)
    for i in range(dim):
        for j in range(dim):
            torch.mul(at[i], bt[j], out=ct[i*dim+j])
    st.finish()
    return c


def outer2D(a, b, c, symmetric):
    """c[2i+j] = a[i]*b[j] (b = a when `symmetric`)."""
    return _outer(a, b, c, symmetric, 2)


def outer3D(a, b, c, symmetric):
    """c[3i+j] = a[i]*b[j] (b = a when `symmetric`)."""
    return _outer(a, b, c, symmetric, 3)


def apply_mask(u_hat, mask):
    """Zero the entries of `u_hat` where `mask` is 0; `mask` broadcasts over the trailing
    dimensions, a leading component dimension of `u_hat` is allowed (applymask.pyx:15-169)."""
    if mask is None:
        return u_hat
    st = Staged()
    ut = st(u_hat, out=True)
    mt = st(mask)
    nd = mt.ndim
    tail = ut.shape[ut.ndim-nd:]
    m = (mt != 0).to(torch.float64).expand(tail).contiguous()
    if ut.dtype != torch.complex128:
        ut.mul_(m)
    else:
        ops.apply_mask(ut, m)
    st.finish()
    return u_hat
