"""Oracle: banded line solvers and the Chebyshev Helmholtz solver, vectorised over lines.

Restates:
  shenfun/la.py:331-393 (TDMA LU/solve), :450-543 (PDMA), :545-613 (FDMA) as a generic
     banded LU without pivoting on diagonals spaced `step` apart;
  shenfun/la.py:1183-1434 (SolverGeneric1ND: A(line) = sum_m scale_m[line] M_m);
  shenfun/chebyshev/la.py:131-351 + optimization/cython/la.pyx:708-947 (Helmholtz LU + solve).
Arrays are [outer][n][inner]; matrices are dense numpy here (sizes the oracle runs are small).
"""
import numpy as np


def banded_lu_solve(A, b, step=1):
    """Solve A x = b (A dense (n, n) or (L, n, n) per line; b (n, L)) by LU without pivoting,
    eliminating in the natural order (what TDMA/PDMA/FDMA do on their diagonals)."""
    A = np.array(A, dtype=float)
    per_line = A.ndim == 3
    n = A.shape[-1]
    x = np.array(b, dtype=complex if np.iscomplexobj(b) else float)
    lo = [k for k in range(1, n) if np.any(np.diagonal(A, -k, -2, -1) != 0)]
    up = [k for k in range(1, n) if np.any(np.diagonal(A, k, -2, -1) != 0)]
    kl = max(lo) if lo else 0
    ku = max(up) if up else 0
    for i in range(n):
        for r in range(i+1, min(n, i+kl+1)):
            l = A[..., r, i]/A[..., i, i]
            if not np.any(l):
                continue
            hi = min(n, i+ku+1)
            if per_line:
                A[:, r, i:hi] -= l[:, None]*A[:, i, i:hi]
                x[r] -= l*x[i]
            else:
                A[r, i:hi] -= l*A[i, i:hi]
                x[r] -= l*x[i]
    for i in range(n-1, -1, -1):
        hi = min(n, i+ku+1)
        if per_line:
            s = np.einsum('lj,jl->l', A[:, i, i+1:hi], x[i+1:hi]) if hi > i+1 else 0
            x[i] = (x[i]-s)/A[:, i, i]
        else:
            s = A[i, i+1:hi] @ x[i+1:hi] if hi > i+1 else 0
            x[i] = (x[i]-s)/A[i, i]
    return x


def solve_axis(A, u, axis):
    """Shared matrix A (n x n) along `axis`; rows beyond n are left untouched."""
    n = A.shape[0]
    sl = [slice(None)]*u.ndim
    sl[axis] = slice(0, n)
    b = np.moveaxis(u[tuple(sl)], axis, 0)
    x = banded_lu_solve(A, b.reshape(n, -1)).reshape(b.shape)
    out = np.array(u)
    out[tuple(sl)] = np.moveaxis(x, 0, axis)
    return out


def solve_generic_1nd(mats, scales, b, axis, pin_zero_line=None):
    """(sum_m scales[m][line] mats[m]) u[line] = b[line]; scales broadcastable arrays with
    extent 1 along `axis` (la.py:1231-1263, 1346-1364).  pin_zero_line=(value,) replaces row 0
    of the line with index 0 by the identity (la.py:355-364, 1352-1360)."""
    n = mats[0].shape[0]
    sl = [slice(None)]*b.ndim
    sl[axis] = slice(0, n)
    bb = np.moveaxis(b[tuple(sl)], axis, 0)
    shp = bb.shape
    bb = bb.reshape(n, -1).copy()
    L = bb.shape[1]
    line_shape = list(b.shape)
    line_shape[axis] = 1
    A = np.zeros((L, n, n))
    for M, s in zip(mats, scales):
        sv = np.broadcast_to(s, line_shape).reshape(-1)
        A += sv[:, None, None]*M[None, :, :]
    if pin_zero_line is not None:
        A[0, 0, :] = 0
        A[0, 0, 0] = 1
        bb[0, 0] = pin_zero_line[0]
    x = banded_lu_solve(A, bb).reshape(shp)
    out = np.array(b, dtype=x.dtype)
    out[tuple(sl)] = np.moveaxis(x, 0, axis)
    return out


def helmholtz_solve(A, B, alfa, beta, b, axis):
    """Chebyshev Dirichlet Helmholtz, line by line as la.pyx:744-756 (LU) and :896-905 (solve).
    A, B: dense (n, n) stiffness / mass; alfa, beta broadcastable with extent 1 along axis."""
    n = A.shape[0]
    sl = [slice(None)]*b.ndim
    sl[axis] = slice(0, n)
    f = np.moveaxis(b[tuple(sl)], axis, 0)
    shp = f.shape
    f = f.reshape(n, -1)
    line_shape = list(b.shape)
    line_shape[axis] = 1
    al = np.broadcast_to(alfa, line_shape).reshape(-1).astype(float)
    be = np.broadcast_to(beta, line_shape).reshape(-1).astype(float)
    A0, A2, A4 = np.diagonal(A), np.diagonal(A, 2), np.diagonal(A, 4)
    Bm2, B0, B2 = np.diagonal(B, -2), np.diagonal(B), np.diagonal(B, 2)
    L = np.zeros((n, len(al)))
    d0 = np.zeros_like(L)
    d1 = np.zeros_like(L)
    d2 = np.zeros_like(L)
    for i in range(n):
        d0[i] = al*A0[i]+be*B0[i]
        if i < n-2:
            d1[i] = al*A2[i]+be*B2[i]
        if i < n-4:
            d2[i] = al*A4[i]
        if i >= 2:
            L[i-2] = be*Bm2[i-2]/d0[i-2]
            d0[i] -= L[i-2]*d1[i-2]
            if i < n-2:
                d1[i] -= L[i-2]*d2[i-2]
            if i < n-4:
                d2[i] -= L[i-2]*d2[i-2]
    y = np.array(f, dtype=complex if np.iscomplexobj(f) else float)
    for i in range(2, n):
        y[i] = y[i]-L[i-2]*y[i-2]
    u = np.zeros_like(y)
    se = np.zeros(y.shape[1], dtype=y.dtype)
    so = np.zeros(y.shape[1], dtype=y.dtype)
    for i in range(n-1, -1, -1):
        v = y[i]
        if i+2 < n:
            v = v-d1[i]*u[i+2]
        if i+4 < n:
            if i % 2 == 0:
                se = se+u[i+4]
                v = v-d2[i]*se
            else:
                so = so+u[i+4]
                v = v-d2[i]*so
        u[i] = v/d0[i]
    out = np.array(b, dtype=u.dtype)
    out[tuple(sl)] = np.moveaxis(u.reshape(shp), 0, axis)
    return out


def biharmonic_solve(S, A, B, a0, alfa, beta, b, axis):
    """Chebyshev biharmonic (a0 S + alfa A + beta B) u = b.  The reference (la.pyx:975-1308) splits
    odd/even unknowns, eliminates with two lower diagonals and keeps three upper diagonals plus a
    rank-2 tail; mathematically this is LU without pivoting of the dense matrix, which is what is
    restated here per line (dense, O(n^3): small sizes only)."""
    n = S.shape[0]
    sl = [slice(None)]*b.ndim
    sl[axis] = slice(0, n)
    f = np.moveaxis(b[tuple(sl)], axis, 0)
    shp = f.shape
    f = f.reshape(n, -1)
    line_shape = list(b.shape)
    line_shape[axis] = 1
    a = np.broadcast_to(a0, line_shape).reshape(-1).astype(float)
    al = np.broadcast_to(alfa, line_shape).reshape(-1).astype(float)
    be = np.broadcast_to(beta, line_shape).reshape(-1).astype(float)
    out_lines = np.zeros(f.shape, dtype=complex if np.iscomplexobj(f) else float)
    for l in range(f.shape[1]):
        H = a[l]*S+al[l]*A+be[l]*B
        x = f[:, l].astype(out_lines.dtype)
        # LU without pivoting (Doolittle, natural order)
        U = H.copy()
        for i in range(n):
            for r in range(i+1, min(n, i+5)):
                if U[r, i] != 0:
                    m = U[r, i]/U[i, i]
                    U[r, i:] -= m*U[i, i:]
                    x[r] -= m*x[i]
        for i in range(n-1, -1, -1):
            x[i] = (x[i]-U[i, i+1:] @ x[i+1:])/U[i, i]
        out_lines[:, l] = x
    out = np.array(b, dtype=out_lines.dtype)
    out[tuple(sl)] = np.moveaxis(out_lines.reshape(shp), 0, axis)
    return out
