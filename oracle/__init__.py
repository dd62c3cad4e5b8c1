"""CPU oracle for the shenfun tensor-product transform-and-solve hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy/scipy restatement of
the reference algorithms (spectralDNS/shenfun v4.3.0; every function cites the
file:line it follows).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker or
the timed CPU baseline -- never from shenfun_b200 (the product path has no CPU
fallback).

Pinning: the reference ships no golden vectors for this path.
The oracle is pinned against outputs of the reference's own Python sources run
in the build container on top of oracle/refshim (serial stand-ins for the
absent third-party mpi4py / mpi4py-fft / FFTW, restated on scipy.fft); those
outputs are committed under tests/golden/ with the generating script
tests/golden/make_golden.py.
"""
