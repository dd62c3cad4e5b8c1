"""Device-resident `Function` (spectral) and `Array` (physical) containers.

Host-side mirror of shenfun/forms/arguments.py:1204-1345 (`ShenfunBaseArray`),
:1427-1888 (`Function`) and :1891-2015 (`Array`).  The reference subclasses
`mpi4py_fft.DistArray` (a numpy ndarray); here the storage is a CUDA
`torch.Tensor` (float64 / complex128) and the classes subclass `torch.Tensor`,
so slicing, in-place arithmetic and broadcasting keep working the way user
code expects, while `.forward()` / `.backward()` call the device plan of the
space they were made for.  Each rank holds its local block only
(`space.local_slice`).
"""
import numbers
import numpy as np
import torch

__all__ = ('Function', 'Array')


def _device():
    if not torch.cuda.is_available():
        # host logic (shapes, slicing) is testable on CPU; kernels are not
        return torch.device('cpu')
    return torch.device('cuda', torch.cuda.current_device())


def _to_tensor(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.as_subclass(torch.Tensor).to(device=device, dtype=dtype)
    if isinstance(x, np.ndarray):
        if np.iscomplexobj(x) and not dtype.is_complex:
            raise TypeError('cannot store complex values in a real Array')
        return torch.from_numpy(np.ascontiguousarray(x)).to(device=device, dtype=dtype)
    return x


def _refine_axis(base, u, ax, n_new):
    """One axis of `Function.refine`: the coefficient array of `base` (N points) along `ax`
    re-expressed for `n_new` points."""
    from .fourier import R2C, C2C
    n_old = base.N
    shape = list(u.shape)
    sl = lambda s: (slice(None),)*ax + (s,)                         # noqa: E731
    if isinstance(base, R2C):
        m_old, m_new = shape[ax], n_new//2+1
        shape[ax] = m_new
        out = u.new_zeros(shape)
        if m_new >= m_old:
            out[sl(slice(0, m_old))] = u
            if n_old % 2 == 0:
                out[sl(m_old-1)] = out[sl(m_old-1)].real*0.5
        else:
            out.copy_(u[sl(slice(0, m_new))])
            if n_old % 2 == 0:
                out[sl(m_new-1)] = out[sl(m_new-1)].real*2
        return out
    if isinstance(base, C2C):
        shape[ax] = n_new
        out = u.new_zeros(shape)
        if n_new >= n_old:
            h = n_old//2
            out[sl(slice(0, h+1))] = u[sl(slice(0, h+1))]
            out[sl(slice(n_new-h, None))] = u[sl(slice(n_old-h, None))]
            if n_old % 2 == 0:
                out[sl(h)] *= 0.5
                out[sl(n_new-h)] *= 0.5
        else:
            h = n_new//2
            out[sl(slice(0, h+1))] = u[sl(slice(0, h+1))]
            out[sl(slice(n_new-h, None))] += u[sl(slice(n_old-h, None))]
        return out
    # polynomial bases: modes first, boundary dofs last
    nbc = base.bcs.num_bcs() if hasattr(base, 'bcs') else 0
    shape[ax] = n_new
    out = u.new_zeros(shape)
    k = min(n_old, n_new)-nbc
    out[sl(slice(0, k))] = u[sl(slice(0, k))]
    if nbc:
        out[sl(slice(n_new-nbc, None))] = u[sl(slice(n_old-nbc, None))]
    return out


class ShenfunBaseArray(torch.Tensor):
    _forward_output = None

    @staticmethod
    def __new__(cls, space, val=0, buffer=None):
        fo = cls._forward_output
        shape = tuple(space.shape(fo))
        dtype = space.spectral_dtype if fo else space.physical_dtype
        dev = _device()
        if buffer is None:
            data = torch.empty(shape, dtype=dtype, device=dev)
            if isinstance(val, (list, tuple)):
                assert len(val) == shape[0]
                for i, v in enumerate(val):
                    data[i] = v
            elif val is not None:           # val=None: left uninitialised (about to be overwritten)
                data.fill_(val)
        elif isinstance(buffer, torch.Tensor):
            data = buffer.as_subclass(torch.Tensor)
            if tuple(data.shape) != shape or data.dtype != dtype:
                raise ValueError(f'buffer has shape {tuple(data.shape)}/{data.dtype}, '
                                 f'space needs {shape}/{dtype}')
            if data.device != dev:
                data = data.to(dev)
        elif isinstance(buffer, np.ndarray):
            if tuple(buffer.shape) != shape:
                raise ValueError(f'buffer has shape {buffer.shape}, space needs {shape}')
            data = _to_tensor(buffer, dtype, dev)
        else:
            # sympy expression(s) evaluated on the local quadrature mesh
            # (forms/arguments.py:1286-1330); a Function is then projected.
            data = cls._from_sympy(space, buffer, shape, dtype, dev)
        obj = torch.Tensor._make_subclass(cls, data)
        obj.__dict__['_space'] = space
        return obj

    @classmethod
    def _from_sympy(cls, space, buffer, shape, dtype, dev):
        import sympy as sp
        exprs = list(buffer) if isinstance(buffer, (list, tuple)) else [buffer]
        pshape = tuple(space.shape(False))
        phys = np.zeros(pshape, dtype=np.float64 if not space.physical_dtype.is_complex else np.complex128)
        mesh = space.local_mesh(True)
        names = 'xyzrs'
        for i, e in enumerate(exprs):
            e = sp.sympify(e)
            syms = sorted(e.free_symbols, key=lambda s: names.index(str(s)) if str(s) in names else 99)
            args = [mesh[names.index(str(s))] for s in syms]
            val = sp.lambdify(syms, e, 'numpy')(*args) if syms else float(e) if e.is_real else complex(e)
            if isinstance(buffer, (list, tuple)):
                phys[i] = val
            else:
                phys[...] = val
        pdata = torch.from_numpy(phys).to(dev)
        if not cls._forward_output:
            return pdata
        arr = Array(space, buffer=pdata)
        out = torch.empty(shape, dtype=dtype, device=dev)
        fun = torch.Tensor._make_subclass(Function, out)
        fun.__dict__['_space'] = space
        space.forward(arr, fun)
        return out

    # -- subclass bookkeeping ---------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        ret = super().__torch_function__(func, types, args, kwargs or {})
        if isinstance(ret, ShenfunBaseArray) and '_space' not in ret.__dict__:
            src = None
            for a in args:
                if isinstance(a, ShenfunBaseArray) and '_space' in a.__dict__:
                    src = a
                    break
            if src is not None and ret.shape == src.shape and ret.dtype == src.dtype:
                ret.__dict__['_space'] = src.__dict__['_space']
            else:
                ret = ret.as_subclass(torch.Tensor)
        return ret

    def __getitem__(self, i):
        space = self.__dict__.get('_space')
        out = torch.Tensor.__getitem__(self.as_subclass(torch.Tensor), i)
        if space is not None and isinstance(i, numbers.Integral) and space.is_composite_space:
            out = out.as_subclass(type(self))
            out.__dict__['_space'] = space[int(i)]
        return out

    def __setitem__(self, i, value):
        base = self.as_subclass(torch.Tensor)
        if isinstance(value, (np.ndarray, torch.Tensor)):
            value = _to_tensor(value, base.dtype, base.device)
        base[i] = value

    # -- reference API --------------------------------------------------------------
    def function_space(self):
        return self.__dict__['_space']

    @property
    def v(self):
        """Plain `torch.Tensor` view of the same memory (arguments.py:1396-1399 `.v`)."""
        return self.as_subclass(torch.Tensor)

    @property
    def rank(self):
        return self.function_space().rank

    @property
    def dimensions(self):
        return self.function_space().dimensions

    @property
    def global_shape(self):
        return self.function_space().global_shape(self._forward_output)

    def local_slice(self):
        return self.function_space().local_slice(self._forward_output)

    @property
    def forward_output(self):
        return self._forward_output

    def get(self):
        """Local block as a host numpy array (device->host copy)."""
        return self.v.detach().cpu().numpy()

    numpy = get

    def copy(self):
        out = self.v.clone().as_subclass(type(self))
        out.__dict__['_space'] = self.function_space()
        return out

    def __reduce_ex__(self, proto):
        raise TypeError('device arrays are not picklable; use .get()')


class Array(ShenfunBaseArray):
    """Values on the (local part of the) quadrature mesh -- arguments.py:1891-2015."""
    _forward_output = False

    def forward(self, output_array=None, kind=None):
        """arguments.py:1968-1983"""
        space = self.function_space()
        if output_array is None:
            output_array = Function(space, val=None)
        return space.forward(self, output_array, kind=kind)

    def offset(self):
        return [s.start for s in self.function_space().local_slice(False)]


class Function(ShenfunBaseArray):
    """Expansion coefficients -- arguments.py:1427-1888."""
    _forward_output = True

    def backward(self, output_array=None, kind=None, mesh=None, padding_factor=None):
        """arguments.py:1579-1638: `padding_factor` selects the dealiased twin space."""
        space = self.function_space()
        if padding_factor is not None:
            space = space.get_dealiased(padding_factor)
        if output_array is None:
            output_array = Array(space, val=None)
        return space.backward(self, output_array, kind=kind)

    def mask_nyquist(self, mask=None):
        """arguments.py:1689: zero the Nyquist coefficients in place."""
        self.function_space().mask_nyquist(self, mask=mask)
        return self

    def assign(self, u_hat):
        """arguments.py:1693-1733: copy self into `u_hat`, a Function on a space made of the same
        bases at other resolutions."""
        space, newspace = self.function_space(), u_hat.function_space()
        if space.is_composite_space:
            for i in range(len(space.flatten())):
                self[i].assign(u_hat[i])
            return u_hat
        for b0, b1 in zip(space.bases, newspace.bases):
            if type(b0) is not type(b1):
                raise AssertionError('Can only assign on spaces with the same underlying bases')
        return self.refine([b.N for b in newspace.bases], output_array=u_hat)

    def refine(self, N, output_array=None):
        """arguments.py:1735-1827: the same expansion on N quadrature points per axis -- zero-padded
        where N grows, truncated where it shrinks.  Per-axis rules as in the reference:
        fourier/bases.py:409-435 (R2C), :570-598 (C2C, symmetric interpolator: an even-N Nyquist
        mode is split over +-N/2 when padding, summed when truncating), spectralbase.py:1363-1390
        (polynomial bases; boundary dofs stay the last rows).  The reference's polynomial
        truncation indexes with the source size and raises; here the low modes are kept.
        On a distributed space the reference redistributes between the axes (`DistArray.redistribute`);
        here every rank gathers the spectral array, refines it and keeps its block of the refined
        space -- refine is set-up / post-processing work, not part of a time step."""
        space = self.function_space()
        if space.is_composite_space:
            comps = [self[i].refine(N) for i in range(len(space.flatten()))]
            if output_array is None:
                from .tensorproductspace import VectorSpace
                output_array = Function(VectorSpace(comps[0].function_space()), val=None)
            for i, c in enumerate(comps):
                output_array[i] = c.v
            return output_array
        nd = len(space.bases)
        if isinstance(N, numbers.Number):
            N = (int(N),)*nd
        if len(N) != nd:
            raise ValueError(f'N has {len(N)} entries, the space has {nd} axes')
        distributed = getattr(space.comm, 'Get_size', lambda: 1)() > 1
        cur, fresh = self.v, False
        if distributed:
            gshape, lshape = tuple(space.global_shape(True)), tuple(self.shape)
            for ax in range(nd):
                if gshape[ax] != lshape[ax]:
                    cur, fresh = space.comm.allgather_blocks(cur, ax), True
        for ax, (base, n_new) in enumerate(zip(space.bases, N)):
            if int(n_new) != base.N:
                cur, fresh = _refine_axis(base, cur, ax, int(n_new)), True
        if distributed:
            newspace = output_array.function_space() if output_array is not None else space.get_refined(N)
            cur = cur[tuple(newspace.local_slice(True))].contiguous()
            if output_array is None:
                return Function(newspace, buffer=cur)
        if output_array is None:
            return Function(space.get_refined(N), buffer=cur if fresh else cur.clone())
        if tuple(output_array.shape) != tuple(cur.shape):
            raise ValueError(f'output_array has shape {tuple(output_array.shape)}, refined array {tuple(cur.shape)}')
        output_array.v.copy_(cur)
        return output_array

    def set_boundary_dofs(self):
        """arguments.py:1525-1539 (no-op for spaces without boundary lifting)."""
        space = self.function_space()
        if hasattr(space, '_set_boundary_dofs'):
            space._set_boundary_dofs(self)
        return self
