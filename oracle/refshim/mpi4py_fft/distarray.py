"""DistArray for a size-1 communicator (test infrastructure only)."""
from numbers import Number
import numpy as np
from mpi4py import MPI
from .pencil import Pencil, Subcomm

comm = MPI.COMM_WORLD


class DistArray(np.ndarray):
    def __new__(cls, global_shape, subcomm=None, val=None, dtype=float,
                buffer=None, strides=None, alignment=None, rank=0):
        if len(global_shape[rank:]) < 2:
            obj = np.ndarray.__new__(cls, global_shape, dtype=dtype, buffer=buffer, strides=strides)
            if buffer is None and isinstance(val, Number):
                obj.fill(val)
            obj._rank = rank
            obj._p0 = None
            return obj
        if isinstance(subcomm, Subcomm):
            pass
        else:
            subcomm = Subcomm(comm, [1]*len(global_shape[rank:]))
        if alignment is None:
            alignment = len(global_shape[rank:]) - 1
        p0 = Pencil(subcomm, global_shape[rank:], axis=alignment)
        subshape = p0.subshape
        if rank > 0:
            subshape = tuple(global_shape[:rank]) + subshape
        obj = np.ndarray.__new__(cls, subshape, dtype=dtype, buffer=buffer)
        if buffer is None and isinstance(val, Number):
            obj.fill(val)
        obj._p0 = p0
        obj._rank = rank
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self._p0 = getattr(obj, '_p0', None)
        self._rank = getattr(obj, '_rank', None)

    @property
    def alignment(self):
        return self._p0.axis

    @property
    def global_shape(self):
        return self.shape[:self.rank] + self._p0.shape

    @property
    def substart(self):
        return (0,)*self.rank + self._p0.substart

    @property
    def subcomm(self):
        return self._p0.subcomm

    @property
    def commsizes(self):
        return [1 for _ in self._p0.subcomm]

    @property
    def pencil(self):
        return self._p0

    @property
    def rank(self):
        return self._rank

    @property
    def dimensions(self):
        return len(self._p0.shape)

    @property
    def v(self):
        return self.__array__()

    def __getitem__(self, i):
        if self.ndim == 1:
            return np.ndarray.__getitem__(self, i)
        if isinstance(i, int) and self.rank > 0:
            v0 = np.ndarray.__getitem__(self, i)
            v0._rank = self.rank - (self.ndim - v0.ndim)
            return v0
        if isinstance(i, tuple) and self.rank == 2:
            v0 = np.ndarray.__getitem__(self, i)
            if isinstance(v0, DistArray):
                v0._rank = self.rank - (self.ndim - v0.ndim)
            return v0
        return np.ndarray.__getitem__(self.v, i)

    def local_slice(self):
        v = [slice(start, start+shape) for start, shape in zip(self._p0.substart, self._p0.subshape)]
        return tuple([slice(0, s) for s in self.shape[:self.rank]] + v)

    def get(self, gslice):
        return np.array(self.v[gslice])

    def redistribute(self, axis=None, out=None):
        if out is None:
            p1 = self._p0.pencil(axis)
            out = DistArray(self.global_shape, subcomm=p1.subcomm, dtype=self.dtype,
                            alignment=axis, rank=self.rank)
            out[...] = self
            return out
        out[...] = self
        return out


def newDistArray(pfft, forward_output=True, val=0, rank=0, view=False):
    global_shape = pfft.global_shape(forward_output)
    p0 = pfft.pencil[forward_output]
    dtype = pfft.forward.output_array.dtype if forward_output else pfft.forward.input_array.dtype
    global_shape = (len(global_shape),)*rank + tuple(global_shape)
    z = DistArray(global_shape, subcomm=p0.subcomm, val=val, dtype=dtype, alignment=p0.axis, rank=rank)
    return z.v if view else z
