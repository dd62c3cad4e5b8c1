"""Size-1 communicator surface of mpi4py.MPI (test infrastructure only)."""
import time
import numpy as np

SUM = 'SUM'
MAX = 'MAX'
MIN = 'MIN'
PROD = 'PROD'
IN_PLACE = 'IN_PLACE'
UNDEFINED = -32766
ORDER_C = 0
DOUBLE = np.dtype('d')
FLOAT = np.dtype('f')
C_DOUBLE_COMPLEX = np.dtype('D')
INT = np.dtype('i')


class Datatype:
    def Create_subarray(self, *a, **k):
        return self

    def Commit(self):
        return self

    def Free(self):
        pass


def _typedict_get(dtype):
    return Datatype()


class _TypeDict(dict):
    def __missing__(self, key):
        return Datatype()


_typedict = _TypeDict()


class Comm:
    def __init__(self, dims=None):
        self.dims = dims

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    size = property(Get_size)
    rank = property(Get_rank)

    def Dup(self):
        return Comm(self.dims)

    def Free(self):
        pass

    def Split(self, color=0, key=0):
        return Comm()

    def Create_cart(self, dims, periods=None, reorder=False):
        return Cartcomm(list(dims))

    def barrier(self):
        pass

    Barrier = barrier

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]

    def allgather(self, obj):
        return [obj]

    def reduce(self, obj, op=SUM, root=0):
        return obj

    def allreduce(self, obj, op=SUM):
        return obj

    def scatter(self, obj, root=0):
        return obj[0]

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        if sendbuf is not IN_PLACE:
            np.asarray(recvbuf)[...] = np.asarray(sendbuf)

    def Reduce(self, sendbuf, recvbuf, op=SUM, root=0):
        self.Allreduce(sendbuf, recvbuf, op)

    def Bcast(self, buf, root=0):
        pass

    def Alltoallw(self, send, recv):
        raise RuntimeError('serial shim: Alltoallw is handled in mpi4py_fft.pencil.Transfer')

    def Abort(self, code=1):
        raise SystemExit(code)


class Cartcomm(Comm):
    def Sub(self, remain_dims):
        return Comm()

    def Get_coords(self, rank):
        return [0]*len(self.dims)

    def Get_topo(self):
        n = len(self.dims)
        return (self.dims, [0]*n, [0]*n)


Intracomm = Comm
COMM_WORLD = Comm()
COMM_SELF = Comm()
COMM_NULL = None


def Compute_dims(nnodes, dims):
    if isinstance(dims, int):
        return [1]*dims
    return [1 if d == 0 else d for d in dims]


def Wtime():
    return time.perf_counter()


def Get_version():
    return (3, 1)
