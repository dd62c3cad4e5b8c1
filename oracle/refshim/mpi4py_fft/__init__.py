"""Serial stand-in for the `mpi4py_fft` package (TEST INFRASTRUCTURE ONLY).

`mpi4py-fft` (and FFTW3 beneath it) is an unpinned third-party dependency of
the reference (/root/reference/setup.py:113, environment.yml:8-19) that is
absent from this image.  The reference's FFT/DCT arithmetic and its global
redistribution live there; this shim restates the *published* behaviour of
that surface (un-normalised FFTW r2c/c2r/c2c and REDFT00/10/01, RODFT00,
`get_normalization() = 1/prod(s)`; Pencil/Subcomm/Transfer for a size-1
communicator where a transfer is a copy) on top of scipy.fft (pocketfft,
fp64).  It exists so the reference's own Python sources can be imported in the
build container to produce golden vectors; it is never on the product path.
"""
from . import fftw  # noqa: F401
from .distarray import DistArray, newDistArray  # noqa: F401
from .mpifft import PFFT, Transform  # noqa: F401
from .pencil import Subcomm, Pencil  # noqa: F401
