"""Subcomm / Pencil / Transfer for a size-1 communicator (test infrastructure)."""
import numpy as np
from mpi4py import MPI


class Subcomm(tuple):
    def __new__(cls, comm, dims=None, reorder=True):
        if isinstance(dims, int):
            ndims = dims
        else:
            ndims = len(dims)
        return super().__new__(cls, [MPI.Comm() for _ in range(ndims)])

    def destroy(self):
        pass


class Transfer:
    """Global redistribution; with one rank it is a plain copy."""

    def __init__(self, comm, shape, dtype, subshapeA, axisA, subshapeB, axisB):
        self.comm = comm
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.subshapeA, self.axisA = tuple(subshapeA), axisA
        self.subshapeB, self.axisB = tuple(subshapeB), axisB

    def forward(self, arrayA, arrayB):
        arrayB[...] = arrayA

    def backward(self, arrayB, arrayA):
        arrayA[...] = arrayB

    def destroy(self):
        pass


class Pencil:
    def __init__(self, subcomm, shape, axis=-1):
        assert len(shape) >= len(subcomm)
        self.shape = tuple(int(i) for i in shape)
        self.axis = axis % len(shape)
        self.subcomm = subcomm
        self.subshape = tuple(self.shape)
        self.substart = (0,)*len(shape)

    def pencil(self, axis):
        return Pencil(self.subcomm, self.shape, axis)

    def transfer(self, pencil, dtype):
        return Transfer(self.subcomm[self.axis], self.shape, dtype,
                        self.subshape, self.axis, pencil.subshape, pencil.axis)
