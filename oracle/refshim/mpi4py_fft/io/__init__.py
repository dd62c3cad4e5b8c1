"""h5py/netCDF4 are absent: file classes are placeholders (I/O is out of scope)."""


class _NoIO:
    def __init__(self, *a, **k):
        raise RuntimeError('serial shim: file I/O not available')


HDF5File = NCFile = _NoIO


def generate_xdmf(*a, **k):
    raise RuntimeError('serial shim: file I/O not available')
