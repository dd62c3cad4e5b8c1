"""Run-time switches (the analogue of shenfun/config.py:9-94).

Only the keys that mean something on the device path are kept.
"""
config = {
    'transforms': {
        # Legendre/Jacobi transforms run as dense fp64 contractions on the
        # device ('vandermonde' definition, spectralbase.py:473-483,539-548)
        'kind': {'legendre': 'vandermonde', 'jacobi': 'vandermonde',
                 'chebyshev': 'fast', 'fourier': 'fast'},
    },
    'fft': {
        'force_generic': False,   # route every FFT through the any-n kernel
    },
}
