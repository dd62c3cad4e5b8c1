"""The pencil half of the mpi4py-fft surface (`mpi4py_fft.pencil.Subcomm / Pencil / Transfer`,
`mpi4py_fft.distarray.DistArray`) as the reference uses it (shenfun/tensorproductspace.py:401,
424-446, 506-510; forms/arguments.py:1204-1345), on top of `transfer.SlabExchange`.

Scope: the SLAB decomposition -- a process grid with ONE distributed axis at a time, which is what
`TensorProductSpace(..., slab=True)` builds and what <= 8 GPUs on NVLink need.  A grid with more
than one distributed axis (pencil decomposition proper) raises NotImplementedError: the device
path has no second exchange.  Arrays are torch tensors on this rank's device.
"""
import numpy as np
import torch

from .comm import comm as default_comm
from .transfer import SlabExchange, _block

__all__ = ('Subcomm', 'Pencil', 'Transfer', 'DistArray')


class _AxisComm:
    """What `subcomm[i]` is used for: size and rank of the process group along one axis."""

    def __init__(self, comm, size, rank):
        self.comm, self._size, self._rank = comm, size, rank

    def Get_size(self):
        return self._size

    def Get_rank(self):
        return self._rank

    size = property(Get_size)
    rank = property(Get_rank)


class Subcomm(tuple):
    """`Subcomm(comm, dims)`: dims[i] = 1 keeps axis i whole, 0 lets it take the ranks
    (mpi4py_fft.pencil.Subcomm; tensorproductspace.py:401).  Exactly one axis may be distributed."""

    def __new__(cls, comm=None, dims=None):
        comm = default_comm if comm is None else comm
        P = comm.Get_size()
        if isinstance(dims, int):
            dims = [0]+[1]*(dims-1)
        dims = list(dims)
        free = [i for i, d in enumerate(dims) if d == 0]
        fixed = int(np.prod([d for d in dims if d > 0])) if any(d > 0 for d in dims) else 1
        if P % fixed != 0:
            raise ValueError('process grid does not match the communicator size')
        rest = P//fixed
        for k, i in enumerate(free):
            dims[i] = rest if k == 0 else 1
        if int(np.prod(dims)) != P:
            raise ValueError('process grid does not match the communicator size')
        if sum(1 for d in dims if d > 1) > 1:
            raise NotImplementedError('pencil decompositions (more than one distributed axis) are not built: '
                                      'the device path is slab-distributed')
        rank = comm.Get_rank()
        obj = super().__new__(cls, [_AxisComm(comm, d, rank if d > 1 else 0) for d in dims])
        obj.comm = comm
        obj.dims = tuple(dims)
        return obj

    def destroy(self):
        pass


class Pencil:
    """Local block of a global array that is whole along `axis` (mpi4py_fft.pencil.Pencil)."""

    def __init__(self, subcomm, shape, axis=-1):
        if not isinstance(subcomm, Subcomm):
            subcomm = Subcomm(default_comm, subcomm)
        nd = len(shape)
        axis = axis % nd
        assert subcomm[axis].Get_size() == 1, 'the aligned axis cannot be distributed'
        self.subcomm = subcomm
        self.shape = tuple(int(s) for s in shape)
        self.axis = axis
        self.subshape, self.substart = [], []
        for i, n in enumerate(self.shape):
            s, m = _block(n, subcomm[i].Get_size(), subcomm[i].Get_rank())
            self.substart.append(s)
            self.subshape.append(m)
        self.subshape, self.substart = tuple(self.subshape), tuple(self.substart)

    @property
    def dist_axis(self):
        d = [i for i, c in enumerate(self.subcomm) if c.Get_size() > 1]
        return d[0] if d else None

    def pencil(self, axis):
        """The pencil aligned in `axis`: the ranks move to the axis this one is aligned in."""
        axis = axis % len(self.shape)
        if axis == self.axis:
            return Pencil(self.subcomm, self.shape, axis)
        dims = list(self.subcomm.dims)
        dims[self.axis], dims[axis] = dims[axis], dims[self.axis]
        return Pencil(Subcomm(self.subcomm.comm, dims), self.shape, axis)

    def transfer(self, pencil, dtype):
        return Transfer(self, pencil, dtype)


class Transfer:
    """Global redistribution between two pencils of the same global shape
    (mpi4py_fft.pencil.Transfer: `MPI_Alltoallw` in the reference, tensorproductspace.py:89,183,252)."""

    def __init__(self, pencilA, pencilB, dtype):
        assert pencilA.shape == pencilB.shape
        self.pencilA, self.pencilB = pencilA, pencilB
        self.subshapeA, self.subshapeB = pencilA.subshape, pencilB.subshape
        self.dtype = dtype
        # A is whole along pencilA.axis and split along pencilB.axis (for one exchange the two swap)
        self._ex = None
        if pencilA.subcomm.comm.Get_size() > 1:
            ga, sa = pencilA.dist_axis, pencilB.dist_axis
            assert ga is not None and sa is not None and ga != sa
            self._ex = SlabExchange(pencilA.subcomm.comm, pencilA.shape, gather_axis=ga, scatter_axis=sa)

    def forward(self, arrayA, arrayB):
        a, b = _plain(arrayA), _plain(arrayB)
        assert tuple(a.shape) == self.subshapeA and tuple(b.shape) == self.subshapeB
        if self._ex is None:
            b.copy_(a)
        else:
            self._ex(a.contiguous(), b, 1, 'forward')
        return arrayB

    def backward(self, arrayB, arrayA):
        a, b = _plain(arrayA), _plain(arrayB)
        if self._ex is None:
            a.copy_(b)
        else:
            self._ex(b.contiguous(), a, 1, 'backward')
        return arrayA

    def destroy(self):
        self._ex = None


def _plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else torch.as_tensor(t)


class DistArray(torch.Tensor):
    """Distributed array: this rank's block of a global array (mpi4py_fft.distarray.DistArray).
    `DistArray(global_shape, subcomm=, val=, dtype=, buffer=, alignment=, rank=)`."""

    @staticmethod
    def __new__(cls, global_shape, subcomm=None, val=None, dtype=float, buffer=None, alignment=None, rank=0,
                device=None):
        global_shape = tuple(int(s) for s in global_shape)
        nd = len(global_shape)-rank
        if isinstance(subcomm, Subcomm):
            sc = subcomm
        elif subcomm is None:
            dims = [0]+[1]*(nd-1) if alignment is None or alignment % nd != 0 else [1, 0]+[1]*(nd-2)
            sc = Subcomm(default_comm, dims)
        else:
            sc = Subcomm(default_comm, subcomm)
        if alignment is None:
            alignment = [i for i, c in enumerate(sc) if c.Get_size() == 1][-1]
        p = Pencil(sc, global_shape[rank:], alignment)
        tdt = {np.dtype('float64'): torch.float64, np.dtype('complex128'): torch.complex128}.get(np.dtype(dtype)) \
            if not isinstance(dtype, torch.dtype) else dtype
        if tdt is None:
            raise TypeError('DistArray: float64 / complex128 only')
        shape = global_shape[:rank]+p.subshape
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
        if buffer is not None:
            data = _plain(buffer).reshape(shape)
        else:
            data = torch.empty(shape, dtype=tdt, device=device)
            if val is not None:
                data.fill_(val)
        obj = torch.Tensor._make_subclass(cls, data)
        obj.__dict__.update(_p0=p, _rank=rank, _gshape=global_shape)
        return obj

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        ret = super().__torch_function__(func, types, args, kwargs or {})
        if isinstance(ret, DistArray) and '_p0' not in ret.__dict__:
            ret = ret.as_subclass(torch.Tensor)
        return ret

    @property
    def v(self):
        return self.as_subclass(torch.Tensor)

    @property
    def pencil(self):
        return self.__dict__['_p0']

    @property
    def subcomm(self):
        return self.pencil.subcomm

    @property
    def commsizes(self):
        return [c.Get_size() for c in self.pencil.subcomm]

    @property
    def alignment(self):
        return self.pencil.axis

    @property
    def global_shape(self):
        return self.__dict__['_gshape']

    @property
    def tensor_rank(self):
        return self.__dict__['_rank']

    def substart(self):
        return self.pencil.substart

    def local_slice(self):
        r = self.tensor_rank
        return (slice(None),)*r+tuple(slice(s, s+n) for s, n in zip(self.pencil.substart, self.pencil.subshape))

    def redistribute(self, axis=None, out=None):
        """Global redistribution so that `axis` becomes the aligned (whole) axis."""
        p0 = self.pencil
        if axis is None:
            assert out is not None
            axis = out.alignment
        axis = axis % len(p0.shape)
        if axis == p0.axis:
            return self
        p1 = p0.pencil(axis)
        r = self.tensor_rank
        if out is None:
            out = DistArray(self.global_shape, subcomm=p1.subcomm, dtype=self.v.dtype, alignment=axis, rank=r,
                            device=self.v.device)
        tr = Transfer(p0, p1, self.v.dtype)
        if r == 0:
            tr.forward(self.v, out.v)
        else:
            src = self.v.reshape((-1,)+p0.subshape)
            dst = out.v.reshape((-1,)+p1.subshape)
            for k in range(src.shape[0]):
                tr.forward(src[k], dst[k])
        return out

    def get(self, gslice):
        """Global slice gathered on every rank (small diagnostics only)."""
        full = torch.zeros(self.global_shape, dtype=self.v.dtype, device=self.v.device)
        full[self.local_slice()] = self.v
        if self.pencil.subcomm.comm.Get_size() > 1:
            import torch.distributed as dist
            buf = torch.view_as_real(full) if full.is_complex() else full
            dist.all_reduce(buf, group=self.pencil.subcomm.comm.group)
        return full[gslice]
