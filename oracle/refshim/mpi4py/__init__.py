"""Serial stand-in for the `mpi4py` package (TEST INFRASTRUCTURE ONLY).

`mpi4py` is an unpinned third-party dependency of the reference
(/root/reference/setup.py:113) that is absent from this image.  This shim
implements the size-1 communicator surface that the reference touches so that
the reference's own Python sources can be imported *in the build container* to
generate golden vectors (tests/golden/make_golden.py).  It never ships in the
product path and is never imported by `shenfun_b200`.
"""
from . import MPI  # noqa: F401
