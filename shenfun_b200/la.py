"""Linear algebra solvers along one axis of device arrays.

Host-side mirror of shenfun/la.py: the banded solvers `TDMA`, `TDMA_O`, `PDMA`,
`FDMA`, `TwoDMA`, `ThreeDMA`, `HeptaDMA`, `DiagMA` (:291-820), `Solve`/`Solver`,
`SolverGeneric1ND` (:1183-1434) and `SolverDiagonal` (:1023-1049); and of
shenfun/chebyshev/la.py `Helmholtz` (:131-351).  All sweeps run in the batched
kernels of csrc/banded.cu (one line per Fourier wavenumber pair, coalesced along the
batch).  LU factors of a single shared matrix are computed on the host (O(N));
per-line factors of `SolverGeneric1ND` / `Helmholtz` are computed on the device
straight from the O(N) diagonals and the per-line scales -- the reference's
O(N0*N1) Python assembly loop (la.py:1248-1263) has no counterpart.
"""
import os
import numpy as np
import torch

from .matrices import SparseMatrix, TPMatrix, get_simplified_tpmatrices
from . import ops

__all__ = ('LinearFormOperator', 'Biharmonic', 'BandedFactor', 'TDMA', 'TDMA_O', 'PDMA', 'FDMA', 'TwoDMA', 'ThreeDMA', 'HeptaDMA', 'DiagMA',
           'Solve', 'Solver', 'SolverGeneric1ND', 'SolverDiagonal', 'Helmholtz', 'tp_matvec')


def _band_of(offsets):
    """(kl, ku, step) of a set of diagonal offsets."""
    offs = [o for o in offsets]
    step = 2 if all(o % 2 == 0 for o in offs) else 1
    kl = max([0]+[-o//step for o in offs if o < 0])
    ku = max([0]+[o//step for o in offs if o > 0])
    return kl, ku, step


_SUPPORTED = {(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2), (2, 4), (1, 0), (2, 0), (2, 1)}


def _fit_band(kl, ku):
    """Smallest supported kernel instantiation that covers (kl, ku)."""
    cands = [(a, b) for (a, b) in _SUPPORTED if a >= kl and b >= ku]
    if not cands:
        raise NotImplementedError(f'banded solver with kl={kl}, ku={ku}')
    return min(cands, key=lambda t: (t[0]+t[1], t[0]))


def _dia_table(mat, kl, ku, step, n):
    """mat[kl+1+ku][n] with entry [d][i] = M[i, i+(d-kl)*step] (scaled), zero outside."""
    tab = np.zeros((kl+1+ku, n))
    for off, v in mat.items():
        assert off % step == 0
        d = off//step+kl
        v = np.real(np.asarray(v))*np.real(mat.scale)
        if off >= 0:
            tab[d, :len(v)] = v[:n]
        else:
            tab[d, -off:-off+len(v)] = v[:n+off]
    return tab


class BandedFactor:
    """LU factors (no pivoting) of ONE banded matrix shared by every line."""

    def __init__(self, kl, ku, step, fac, n):
        self.kl, self.ku, self.step, self.n = kl, ku, step, n
        self.fac_host = fac
        self._dev = None

    @classmethod
    def from_matrix(cls, mat, constraint_row=None):
        n = mat.shape[0]
        kl, ku, step = _band_of(mat.keys())
        kl, ku = _fit_band(kl, ku)
        A = _dia_table(mat, kl, ku, step, n)
        if constraint_row is not None:
            # la.py:355-364: row replaced by the identity
            A[:, constraint_row] = 0
            A[kl, constraint_row] = 1
        return cls(kl, ku, step, cls._lu(A, kl, ku, step, n), n)

    @staticmethod
    def _lu(A, kl, ku, step, n):
        """Row-wise elimination in the same order as the device kernel
        (banded_factor_kernel): L_p[i], D[i], U_q[i] in table layout."""
        F = A.copy()
        for i in range(n):
            for p in range(kl, 0, -1):
                r = i-p*step
                if r < 0:
                    F[kl-p, i] = 0.0
                    continue
                l = F[kl-p, i]/F[kl, r]
                F[kl-p, i] = l
                for q in range(1, ku+1):
                    col = kl-p+q
                    if col < kl+1+ku:
                        F[col, i] -= l*F[kl+q, r]
        return F

    def device(self):
        if self._dev is None:
            self._dev = ops.as_device_tensor(np.ascontiguousarray(self.fac_host))
        return self._dev

    def solve(self, u, outer, inner, n_u):
        ops.banded_solve(self.kl, self.ku, self.step, u, self.device(), outer, inner, self.n, n_u, True)
        return u


def _axis_view(u, axis):
    shape = u.shape
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis+1:], dtype=np.int64))
    return outer, shape[axis], inner


class BandedMatrixSolver:
    """Base of the named solvers (la.py:183-289): LU once, solve along any axis."""

    def __init__(self, mat):
        if isinstance(mat, TPMatrix):
            nd = [m for m in mat.mats if not m.is_diagonal()]
            assert len(nd) == 1
            mat = nd[0]
        assert isinstance(mat, SparseMatrix)
        self.mat = mat
        self._factor = {}

    def _get_factor(self, constraints):
        key = tuple(c[0] for c in constraints)
        if key not in self._factor:
            row = constraints[0][0] if constraints else None
            self._factor[key] = BandedFactor.from_matrix(self.mat, constraint_row=row)
        return self._factor[key]

    def perform_lu(self):
        return self._get_factor(())

    def __call__(self, b, u=None, axis=0, constraints=()):
        bt = ops.as_device_tensor(b)
        if u is None:
            u = b
            ut = bt
        else:
            ut = ops.plain(u)
            if ut.data_ptr() != bt.data_ptr():
                ut.copy_(bt)
        assert ut.is_contiguous()
        outer, n_u, inner = _axis_view(ut, axis)
        for (row, val) in constraints:
            ut.view(outer, n_u, inner)[:, row, :] = val
        self._get_factor(constraints).solve(ut, outer, inner, n_u)
        if hasattr(u, 'set_boundary_dofs') and getattr(u, 'function_space', None):
            pass
        return u

    solve = __call__


class TDMA(BandedMatrixSolver):
    """Tridiagonal, offsets (-2, 0, 2) (la.py:331-393)."""


class TDMA_O(BandedMatrixSolver):
    """Tridiagonal, offsets (-1, 0, 1) (la.py:395-448)."""


class PDMA(BandedMatrixSolver):
    """Pentadiagonal, offsets (-4, -2, 0, 2, 4) (la.py:450-543)."""


class FDMA(BandedMatrixSolver):
    """Four diagonals, offsets (-2, 0, 2, 4) (la.py:545-613)."""


class TwoDMA(BandedMatrixSolver):
    """Offsets (0, 2) (la.py:615-660)."""


class ThreeDMA(BandedMatrixSolver):
    """Offsets (0, 2, 4) (la.py:662-709)."""


class HeptaDMA(BandedMatrixSolver):
    """Offsets (-4, -2, 0, 2, 4, 6, 8) (la.py:711-820)."""


class DiagMA(BandedMatrixSolver):
    """Diagonal (la.py:291-329)."""


Solve = Solver = BandedMatrixSolver


class SolverDiagonal:
    """All-diagonal (Fourier) tensor-product systems (la.py:1023-1049)."""

    def __init__(self, mats):
        if isinstance(mats, TPMatrix):
            mats = [mats]
        self.mat = get_simplified_tpmatrices(mats)[0]
        assert self.mat.is_diagonal()

    def __call__(self, b, u=None, constraints=()):
        u = self.mat.solve(b, u)
        for (row, val) in constraints:
            sp = self.mat.space
            if all(s.start == 0 for s in sp.local_slice(True)):
                ops.plain(u).view(-1)[row] = val
        return u


def _line_scales(tpmat, axis):
    """Per-line scale of a TPMatrix: product of the diagonal 1-D matrices of the other axes
    (local block), times the scalar scale, flattened over (outer, inner)."""
    space = tpmat.space
    sl = space.local_slice(True)
    nd = tpmat.dimensions
    arr = np.ones([1]*nd)
    for ax, m in enumerate(tpmat.mats):
        if ax == axis:
            continue
        assert m.is_diagonal(), 'only one non-diagonal axis per TPMatrix'
        v = m.scaled_diagonal(0)
        full = np.zeros(space.bases[ax].spectral_len(), dtype=v.dtype)
        full[space.bases[ax].slice()] = v
        shape = [1]*nd
        shape[ax] = sl[ax].stop-sl[ax].start
        arr = arr*full[sl[ax]].reshape(shape)
    arr = arr*tpmat.scale
    if np.iscomplexobj(arr):
        if np.abs(arr.imag).max() > 1e-14*max(1.0, np.abs(arr.real).max()):
            raise NotImplementedError('complex scale arrays (odd total derivative order along Fourier axes)')
        arr = arr.real
    shape = list(space.shape(True))
    shape[axis] = 1
    return np.ascontiguousarray(np.broadcast_to(arr, shape)).reshape(-1)


def _line_scales_complex(tpmat, axis):
    """Like `_line_scales` but returns (real, imag) parts (imag None when it vanishes)."""
    space = tpmat.space
    sl = space.local_slice(True)
    nd = tpmat.dimensions
    arr = np.ones([1]*nd, dtype=complex)
    for ax, m in enumerate(tpmat.mats):
        if ax == axis:
            continue
        v = m.scaled_diagonal(0)
        full = np.zeros(space.bases[ax].spectral_len(), dtype=complex)
        full[space.bases[ax].slice()] = v
        shape = [1]*nd
        shape[ax] = sl[ax].stop-sl[ax].start
        arr = arr*full[sl[ax]].reshape(shape)
    arr = arr*tpmat.scale
    shape = list(space.shape(True))
    if axis is not None:
        shape[axis] = 1
    arr = np.ascontiguousarray(np.broadcast_to(arr, shape)).reshape(-1)
    re = np.ascontiguousarray(arr.real)
    im = np.ascontiguousarray(arr.imag) if np.abs(arr.imag).max() > 0 else None
    if im is not None and np.abs(re).max() == 0:
        re = None
    return re, im


class LinearFormOperator:
    """out (+)= sum_m scale_m(line) M_m u for the TPMatrices of one linear form (same source and
    target component): banded 1-D matrices go through `shb_banded_matvec`, the dense upper
    triangular Chebyshev matrices through the structured kernels (constant rows:
    `shb_helmholtz_matvec`; rank-2 tail: `shb_biharmonic_matvec`), anything else through the
    tensor-core contraction.  Fourier scales may be complex (odd derivative orders)."""

    MAXDIAG = 16

    def __init__(self, mats):
        self.mats = list(mats)
        m0 = self.mats[0]
        self.space = m0.space             # test space
        self.trialspace = m0.trialspace
        nd_axes = set()
        for m in self.mats:
            for ax, mm in enumerate(m.mats):
                if m.space.bases[ax].family() != 'fourier':
                    nd_axes.add(ax)
        if len(nd_axes) > 1:
            raise NotImplementedError('linear forms with more than one non-periodic axis')
        self.axis = nd_axes.pop() if nd_axes else None
        shape_out = self.space.shape(True)
        shape_in = self.trialspace.shape(True)
        self.parts = []
        if self.axis is None:
            d = None
            for m in self.mats:
                a = m.diagonal_array()
                d = a if d is None else d+a
            self.parts.append(('diag', d))
            return
        ax = self.axis
        self.outer = int(np.prod(shape_out[:ax], dtype=np.int64))
        self.inner = int(np.prod(shape_out[ax+1:], dtype=np.int64))
        self.n_out, self.n_u = shape_out[ax], shape_in[ax]
        groups = {}
        for m in self.mats:
            m1 = m.mats[ax]
            sig = (m1.shape, tuple(sorted(m1.keys())))+tuple(np.asarray(m1[k]*m1.scale).tobytes() for k in sorted(m1.keys()))
            re, im = _line_scales_complex(m, ax)
            g = groups.setdefault(sig, [m1, None, None])
            if re is not None:
                g[1] = re if g[1] is None else g[1]+re
            if im is not None:
                g[2] = im if g[2] is None else g[2]+im
        banded = {False: [], True: []}     # keyed by "imaginary scales"
        for m1, re, im in groups.values():
            wide = len(m1.keys()) > self.MAXDIAG
            for imag, sc in ((False, re), (True, im)):
                if sc is None:
                    continue
                scd = ops.as_device_tensor(sc)
                if wide:
                    self.parts.append(self._wide_part(m1, scd, imag))
                else:
                    banded[imag].append((m1, scd))
        for imag, lst in banded.items():
            for i in range(0, len(lst), 4):
                chunk = lst[i:i+4]
                offs = sorted(set(o for m1, _ in chunk for o in m1.keys()))
                if len(offs) > self.MAXDIAG:
                    for m1, scd in chunk:
                        self.parts.append(self._banded_part([(m1, scd)], imag))
                else:
                    self.parts.append(self._banded_part(chunk, imag))

    def _banded_part(self, chunk, imag):
        offs = sorted(set(o for m1, _ in chunk for o in m1.keys()))
        n_rows, n_cols = chunk[0][0].shape
        tabs = []
        for m1, _ in chunk:
            assert m1.shape == (n_rows, n_cols)
            t = np.zeros((len(offs), n_rows))
            for k, off in enumerate(offs):
                if off in m1:
                    v = np.real(m1[off])*np.real(m1.scale)
                    if off >= 0:
                        t[k, :len(v)] = v
                    else:
                        t[k, -off:-off+len(v)] = v
            tabs.append(ops.as_device_tensor(t))
        return ('banded', offs, tabs, [s for _, s in chunk], n_rows, n_cols, imag)

    def _wide_part(self, m1, scd, imag):
        n = m1.shape[0]
        sc = float(np.real(m1.scale))
        offs = sorted(m1.keys())
        if m1.shape[0] != m1.shape[1]:
            # cross-basis matrices with a dense upper part (e.g. (phi_j', T_k)_w): running sums when the
            # tail is separable, else the tensor-core contraction
            st = self._septail_part(m1, scd, imag)
            if st is not None:
                return st
            return ('dense', ops.as_device_tensor(np.ascontiguousarray(m1.diags())), scd, n, imag)

        def diag(off):
            v = np.zeros(n)
            if off in m1:
                d = np.real(m1[off])*sc
                v[:len(d)] = d
            return v
        zeros = ops.as_device_tensor(np.zeros(n))
        if min(offs) >= 0 and all(o % 2 == 0 for o in offs):
            # constant rows beyond +4?  (Chebyshev Dirichlet stiffness)
            if all(np.allclose(m1[o], m1[4][:len(m1[o])], rtol=1e-11, atol=0) for o in offs if o > 4):
                tabs = [ops.as_device_tensor(diag(0)), ops.as_device_tensor(diag(2)), ops.as_device_tensor(diag(4)),
                        zeros, zeros, zeros]
                return ('helm', tabs, scd, zeros.new_zeros(scd.shape), n, imag)
            # rank-2 tail beyond +4?  (Chebyshev biharmonic fourth-derivative matrix)
            k = np.arange(n, dtype=float)
            al = 8*np.pi*k*(k+1)*(k+2)*(k+4)*sc
            be = 24*np.pi*(k+1)*(k+2)*sc
            ok = True
            for o in offs:
                if o >= 6:
                    j = np.arange(len(m1[o]))+o
                    ok = ok and np.allclose(m1[o]*sc, (al[:len(j)]+be[:len(j)]*(j+2.0)**2)/(j+3.0), rtol=1e-10)
            if ok:
                tab = np.zeros((13, n))
                tab[0], tab[1], tab[2], tab[3], tab[4] = diag(0), diag(2), diag(4), al, be
                return ('bih', ops.as_device_tensor(tab), scd, zeros.new_zeros(scd.shape), n, imag)
        st = self._septail_part(m1, scd, imag)
        if st is not None:
            return st
        return ('dense', ops.as_device_tensor(np.ascontiguousarray(m1.diags())), scd, n, imag)

    def _septail_part(self, m1, scd, imag):
        found = separable_tail(m1.diags())
        if found is None:
            return None
        w, r, t, band = found
        if len(band) > self.MAXDIAG:
            return None
        bpart = None
        if band:
            bpart = self._banded_part([(SparseMatrix(band, m1.shape, 1.0), scd)], imag)
        return ('septail', bpart, int(w), ops.as_device_tensor(r), ops.as_device_tensor(t), scd,
                m1.shape[0], m1.shape[1], imag)

    def matvec(self, u, out, accumulate=False):
        """out (+)= sum of the parts applied to u.  Without `accumulate` the first part overwrites
        `out` (no separate zero fill, no temporary)."""
        u, out = ops.plain(u), ops.plain(out)
        acc = bool(accumulate)
        if not self.parts and not acc:
            out.zero_()
        for part in self.parts:
            kind = part[0]
            if kind == 'diag':
                d = part[1]
                d = d.to(u.dtype) if d.dtype != u.dtype else d
                if acc:
                    out.addcmul_(u, d)
                else:
                    torch.mul(u, d, out=out)
            elif kind == 'banded':
                _, offs, tabs, scales, n_rows, n_cols, imag = part
                ops.banded_matvec(u, out, offs, tabs, scales, self.outer, self.inner, n_rows, n_cols,
                                  self.n_u, self.n_out, (1 if acc else 0) | (2 if imag else 0))
            elif kind == 'septail':
                _, bpart, w, r, t, sc, n_rows, n_cols, imag = part
                if bpart is not None:
                    _, offs, tabs, scales, _, _, _ = bpart
                    ops.banded_matvec(u, out, offs, tabs, scales, self.outer, self.inner, n_rows, n_cols,
                                      self.n_u, self.n_out, (1 if acc else 0) | (2 if imag else 0))
                    acc = True
                ops.tail_matvec(u, out, r, t, sc, w, self.outer, self.inner, n_rows, n_cols, self.n_u, self.n_out,
                                (1 if acc else 0) | (2 if imag else 0))
            else:
                tmp = torch.empty_like(out) if acc else out
                if kind == 'helm':
                    _, tabs, sc, z, n, imag = part
                    ops.helmholtz_matvec(u, tmp, tabs, sc, z, self.outer, self.inner, n, self.n_u)
                elif kind == 'bih':
                    _, tab, sc, z, n, imag = part
                    ops.biharmonic_matvec(u, tmp, tab, sc, z, z, self.outer, self.inner, n, self.n_u)
                else:
                    _, M, sc, n, imag = part
                    if not hasattr(self, '_dense_cache'):
                        self._dense_cache = {}
                    full = self._dense_cache.get(id(M))
                    if full is None:
                        full = torch.zeros((self.n_out, self.n_u), dtype=torch.float64, device=u.device)
                        full[:M.shape[0], :M.shape[1]] = M
                        self._dense_cache[id(M)] = full
                    ops.gemm_axis(full, u, tmp, self.outer, self.inner)
                    shape = [self.outer, 1, self.inner]
                    tmp.view(self.outer, self.n_out, self.inner).mul_(sc.view(shape))
                if acc:
                    out.add_(tmp*1j if imag else tmp)
                elif imag:
                    out.mul_(1j)
            acc = True
        return out



def separable_tail(D, rtol=1e-12, min_tail=3):
    """Structure of an upper-triangular derivative matrix: find the smallest offset w such that every
    entry with j - k >= w on the parity class of w is r_k * t_j with r == 1 (column function) or
    t == 1 (row function) and the other parity class is empty beyond w.  Returns (w, r, t, band) with
    `band` = {offset: diagonal} for the offsets below w, or None when there is no such structure
    (or the tail is shorter than `min_tail` diagonals and plain banded storage is just as good)."""
    D = np.asarray(D, dtype=float)
    n_rows, n_cols = D.shape
    K, J = np.indices(D.shape)
    off = J-K
    amax = np.abs(D).max() if D.size else 0.0
    if amax == 0:
        return None
    tol = rtol*amax
    present = sorted(set(off[np.abs(D) > tol].tolist()))
    for w in present:
        if len([o for o in present if o >= w]) < min_tail:
            break
        other = (off > w) & ((off-w) % 2 == 1)
        if np.any(np.abs(D[other]) > tol):
            continue
        tail = (off >= w) & ((off-w) % 2 == 0)
        Dt = np.where(tail, D, np.nan)
        r = t = None
        with np.errstate(all='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', RuntimeWarning)
                cmin, cmax = np.nanmin(Dt, axis=0), np.nanmax(Dt, axis=0)
                rmin, rmax = np.nanmin(Dt, axis=1), np.nanmax(Dt, axis=1)
        if np.all(np.isnan(cmax) | (cmax-cmin <= tol)):
            t, r = np.nan_to_num(cmax), np.ones(n_rows)
        elif np.all(np.isnan(rmax) | (rmax-rmin <= tol)):
            r, t = np.nan_to_num(rmax), np.ones(n_cols)
        if r is None:
            continue
        band = {}
        for o in present:
            if o < w:
                band[o] = np.diagonal(D, o).copy()
        return w, r, t, band
    return None


def _nondiagonal_axis(mats):
    axes = set()
    for m in mats:
        for ax, mm in enumerate(m.mats):
            if not mm.is_diagonal() or m.space.bases[ax].family() != 'fourier':
                axes.add(ax)
    if len(axes) != 1:
        raise NotImplementedError(f'exactly one non-periodic axis is supported, got {sorted(axes)}')
    return axes.pop()


class _LineOperator:
    """sum_m scale_m[line] M_m along `axis`: shared tables for factor / matvec kernels."""

    def __init__(self, mats):
        if isinstance(mats, TPMatrix):
            mats = [mats]
        mats = list(mats)
        self.tpmats = mats
        self.space = mats[0].space
        self.axis = _nondiagonal_axis(mats)
        # merge terms that share the same 1-D matrix along the solve axis: their per-line
        # scales add up (what get_simplified_tpmatrices does, matrixbase.py:1648-1677)
        groups = {}
        for m in mats:
            m1 = m.mats[self.axis]
            sig = (tuple(sorted(m1.keys())),)+tuple(np.asarray(m1[k]*m1.scale).tobytes() for k in sorted(m1.keys()))
            sc = _line_scales(m, self.axis)
            if sig in groups:
                groups[sig][1] += sc
            else:
                groups[sig] = [m1, sc.copy()]
        assert 1 <= len(groups) <= 4, 'up to four distinct 1-D matrices'
        self.mats1d = [g[0] for g in groups.values()]
        self.n_rows, self.n_cols = self.mats1d[0].shape
        self.scales = [ops.as_device_tensor(np.ascontiguousarray(g[1])) for g in groups.values()]
        shape = self.space.shape(True)
        self.outer = int(np.prod(shape[:self.axis], dtype=np.int64))
        self.inner = int(np.prod(shape[self.axis+1:], dtype=np.int64))
        self.n_u = shape[self.axis]
        offs = sorted(set(o for m in self.mats1d for o in m.keys()))
        self.offsets = offs
        self.kl, self.ku, self.step = _band_of(offs)

    def matvec(self, u, out, accumulate=False):
        if not hasattr(self, '_mv_tabs'):
            tabs = []
            for m in self.mats1d:
                t = np.zeros((len(self.offsets), self.n_rows))
                for k, off in enumerate(self.offsets):
                    if off in m:
                        v = np.real(m[off])*np.real(m.scale)
                        if off >= 0:
                            t[k, :len(v)] = v
                        else:
                            t[k, -off:-off+len(v)] = v
                tabs.append(ops.as_device_tensor(t))
            self._mv_tabs = tabs
        ops.banded_matvec(u, out, self.offsets, self._mv_tabs, self.scales, self.outer, self.inner,
                          self.n_rows, self.n_cols, self.n_u, self.n_u, accumulate)
        return out


def tp_matvec(tpmat, u, out, accumulate=False):
    if not hasattr(tpmat, '_lineop'):
        tpmat._lineop = LinearFormOperator([tpmat])
    return tpmat._lineop.matvec(ops.plain(u), ops.plain(out), accumulate)


class SolverGeneric1ND:
    """One banded solve per Fourier wavenumber pair (la.py:1183-1434)."""

    def __init__(self, mats, fixed_gauge=None):
        if isinstance(mats, TPMatrix):
            mats = [mats]
        self.op = _LineOperator(mats)
        self.T = self.op.space
        self._fac = {}

    def _zero_line(self):
        """Index of the line with all-zero Fourier wavenumbers on this rank, or -1
        (la.py:1352-1360 tests the local slice start, not the rank)."""
        sl = self.T.local_slice(True)
        idx = 0
        stride = 1
        for ax in reversed(range(len(sl))):
            if ax == self.op.axis:
                continue
            if sl[ax].start != 0:
                return -1
            stride *= (sl[ax].stop-sl[ax].start)
        return 0

    def _factor(self, pinned):
        if pinned not in self._fac:
            op = self.op
            kl, ku = _fit_band(op.kl, op.ku)
            tabs = [ops.as_device_tensor(_dia_table(m, kl, ku, op.step, op.n_rows)) for m in op.mats1d]
            fac = torch.empty((kl+1+ku, op.outer, op.n_rows, op.inner), dtype=torch.float64,
                              device=tabs[0].device)
            pin_line = self._zero_line() if pinned else -1
            ops.banded_factor(kl, ku, op.step, tabs, op.scales, fac, op.outer, op.inner, op.n_rows, pin_line)
            self._fac[pinned] = (kl, ku, fac)
        return self._fac[pinned]

    def matvec(self, u, c):
        self.op.matvec(ops.plain(u), ops.plain(c), accumulate=False)
        return c

    def __call__(self, b, u=None, constraints=()):
        op = self.op
        bt = ops.as_device_tensor(b)
        if u is None:
            u, ut = b, bt
        else:
            ut = ops.plain(u)
            if ut.data_ptr() != bt.data_ptr():
                ut.copy_(bt)
        pinned = len(constraints) > 0
        if pinned:
            assert len(constraints) == 1 and constraints[0][0] == 0, 'only the (0, value) gauge is supported'
            zl = self._zero_line()
            if zl >= 0:
                v = ut.view(op.outer, op.n_u, op.inner)
                v[zl//op.inner, 0, zl % op.inner] = constraints[0][1]
        kl, ku, fac = self._factor(pinned)
        ops.banded_solve(kl, ku, op.step, ut, fac, op.outer, op.inner, op.n_rows, op.n_u, False)
        if hasattr(u, 'set_boundary_dofs'):
            u.set_boundary_dofs()
        return u


class Helmholtz:
    """Chebyshev-Galerkin Helmholtz solver alfa*A + beta*B (chebyshev/la.py:131-351) for the
    Dirichlet basis: A = (phi_j'', phi_k)_w is upper triangular with constant rows."""

    def __init__(self, *args):
        if len(args) == 1:
            args = args[0]
        mats = list(args)
        assert all(isinstance(m, TPMatrix) for m in mats)
        op = _LineOperator(mats)
        assert len(op.mats1d) == 2, 'Helmholtz needs one stiffness and one mass matrix'
        self.op = op
        self.T = op.space
        base = self.T.bases[op.axis]
        assert base.family() == 'chebyshev' and base.boundary_condition() == 'Dirichlet', \
            'Helmholtz: Chebyshev Dirichlet basis only on the device path'
        # which is the stiffness matrix: the one with the widest upper band
        i_a = int(np.argmax([max(m.keys()) for m in op.mats1d]))
        A, B = op.mats1d[i_a], op.mats1d[1-i_a]
        self.alfa, self.beta = op.scales[i_a], op.scales[1-i_a]
        n = op.n_rows

        def diag(m, off, length):
            v = np.zeros(n)
            if off in m:
                d = np.real(m[off])*np.real(m.scale)
                v[:len(d)] = d
            return ops.as_device_tensor(v)
        assert set(B.keys()) <= {-2, 0, 2}
        # constant rows beyond the +4 diagonal (chebyshev/matrices.py ASDSDmat)
        for off in A.keys():
            if off > 4:
                assert np.allclose(A[off], A[4][:len(A[off])], rtol=1e-12, atol=0), 'A must have constant rows'
        self._tabs = [diag(A, 0, n), diag(A, 2, n-2), diag(A, 4, n-4), diag(B, -2, n-2), diag(B, 0, n), diag(B, 2, n-2)]
        # one stored array per line (1/d0); d1, d2, L are regenerated inside the chunk-parallel solve
        # (csrc/linesolve.cu).  Lines too long for one CTA keep the reference's four arrays.
        self.lean = n <= Helmholtz.LEAN_MAX_N and os.environ.get('SHB_HELMHOLTZ_LEAN', '1') == '1'
        if self.lean:
            self.fac = torch.empty((op.outer, n, op.inner), dtype=torch.float64, device=self.alfa.device)
            ops.helmholtz_factor_lean(self._tabs, self.alfa, self.beta, self.fac, op.outer, op.inner, n)
            self._lean_tabs = ops.helmholtz_lean_tables(self._tabs, n)
        else:
            self.fac = torch.empty((4, op.outer, n, op.inner), dtype=torch.float64, device=self.alfa.device)
            ops.helmholtz_factor(*self._tabs, self.alfa, self.beta, self.fac, op.outer, op.inner, n)

    LEAN_MAX_N = 1024

    def matvec(self, v, c):
        op = self.op
        ops.helmholtz_matvec(ops.plain(v), ops.plain(c), self._tabs, self.alfa, self.beta, op.outer, op.inner,
                             op.n_rows, op.n_u)
        return c

    def __call__(self, b, u=None, constraints=()):
        op = self.op
        bt = ops.as_device_tensor(b)
        if u is None:
            u = b
        ut = ops.plain(u)
        if self.lean:
            ops.helmholtz_solve_lean(bt, ut, self._lean_tabs, self.alfa, self.beta, self.fac, op.outer, op.inner,
                                     op.n_rows, op.n_u)
        else:
            ops.helmholtz_solve(bt, ut, self.fac, op.outer, op.inner, op.n_rows, op.n_u)
        if hasattr(u, 'set_boundary_dofs'):
            u.set_boundary_dofs()
        return u


class Biharmonic:
    """Chebyshev-Galerkin biharmonic solver a0*S + alfa*A + beta*B on the ShenBiharmonic basis
    (chebyshev/la.py:354-558).  The three 1-D matrices are recognised by their shape: S (fourth
    derivative) is upper triangular with a rank-2 tail, A has offsets (-2, 0, 2), B (-4 .. 4)."""

    def __init__(self, *args):
        if len(args) == 1:
            args = args[0]
        mats = list(args)
        assert all(isinstance(m, TPMatrix) for m in mats)
        op = _LineOperator(mats)
        assert len(op.mats1d) == 3, 'Biharmonic needs S, A and B'
        self.op = op
        self.T = op.space
        self.tpmats = mats
        base = self.T.bases[op.axis]
        assert base.family() == 'chebyshev' and base.boundary_condition() == 'Biharmonic', \
            'Biharmonic: Chebyshev ShenBiharmonic basis only'
        order = sorted(range(3), key=lambda i: (max(op.mats1d[i].keys()), -min(op.mats1d[i].keys())))
        iA, iB, iS = order      # A: max offset 2, B: 4, S: widest
        S, A, B = op.mats1d[iS], op.mats1d[iA], op.mats1d[iB]
        assert set(A.keys()) <= {-2, 0, 2} and set(B.keys()) <= {-4, -2, 0, 2, 4} and max(S.keys()) > 4
        self.a0, self.alfa, self.beta = op.scales[iS], op.scales[iA], op.scales[iB]
        n = op.n_rows
        k = np.arange(n, dtype=float)
        tab = np.zeros((13, n))

        def put(row, m, off):
            if off in m:
                d = np.real(m[off])*np.real(m.scale)
                if off >= 0:
                    tab[row, :len(d)] = d
                else:
                    tab[row, -off:-off+len(d)] = d
        put(0, S, 0)
        put(1, S, 2)
        put(2, S, 4)
        # rank-2 tail of S (chebyshev/matrices.py SSBSBmat; la.pyx:1218-1236):
        # S[k, j] = (alpha_k + beta_k (j+2)^2)/(j+3), j >= k+6
        sS = float(np.real(S.scale))
        tab[3] = 8*np.pi*k*(k+1)*(k+2)*(k+4)*sS
        tab[4] = 24*np.pi*(k+1)*(k+2)*sS
        for off in S.keys():
            if off >= 6:
                j = np.arange(len(S[off]))+off
                ref = (tab[3, :len(j)]+tab[4, :len(j)]*(j+2.0)**2)/(j+3.0)
                assert np.allclose(S[off]*sS, ref, rtol=1e-10), 'S does not have the expected tail'
        put(5, A, -2)
        put(6, A, 0)
        put(7, A, 2)
        put(8, B, -4)
        put(9, B, -2)
        put(10, B, 0)
        put(11, B, 2)
        put(12, B, 4)
        self._tab = ops.as_device_tensor(tab)
        self.fac = torch.empty((7, op.outer, n, op.inner), dtype=torch.float64, device=self._tab.device)
        ops.biharmonic_factor(self._tab, self.a0, self.alfa, self.beta, self.fac, op.outer, op.inner, n)

    def matvec(self, v, c):
        op = self.op
        ops.biharmonic_matvec(ops.plain(v), ops.plain(c), self._tab, self.a0, self.alfa, self.beta, op.outer,
                              op.inner, op.n_rows, op.n_u)
        return c

    def __call__(self, b, u=None, constraints=()):
        op = self.op
        bt = ops.as_device_tensor(b)
        if u is None:
            u = b
        ops.biharmonic_solve(bt, ops.plain(u), self.fac, op.outer, op.inner, op.n_rows, op.n_u)
        return u


def mass_solve(T, rhs, out):
    """Solve the (tensor-product) mass system for `project` on non-Fourier spaces."""
    rt, ot = ops.plain(rhs), ops.plain(out)
    if ot.data_ptr() != rt.data_ptr():
        ot.copy_(rt)
    flat = T.flatten() if T.is_composite_space else [T]
    views = [ot[i] for i in range(len(flat))] if T.is_composite_space else [ot]
    for sp, v in zip(flat, views):
        for ax, b in enumerate(sp.bases):
            if b.family() == 'fourier':
                continue
            shape = v.shape
            outer = int(np.prod(shape[:ax], dtype=np.int64))
            inner = int(np.prod(shape[ax+1:], dtype=np.int64))
            b.apply_inverse_mass(v, outer, inner)
    return out
