"""The reference's OWN compiled kernels (shenfun/optimization/cython/{la,Matvec,transforms}.pyx,
cythonised unchanged by oracle/build_ref.sh into oracle/_ref/refcy/): test infrastructure and the
`cpu_baseline.kind = "reference"` legs of bench.py -- never imported by the product path.

    from oracle import refcy
    la = refcy.load('la')        # TDMA_LU, TDMA_Solve, LU_Helmholtz, Solve_Helmholtz, ...
"""
import importlib
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'refcy')


def available():
    return os.path.isdir(_DIR) and any(f.startswith('la.') and f.endswith('.so') for f in os.listdir(_DIR))


def load(name):
    if not available():
        raise ImportError('oracle/_ref/refcy is not built (run oracle/build_ref.sh where /root/reference exists)')
    if _DIR not in sys.path:
        sys.path.insert(0, _DIR)
    return importlib.import_module(name)
