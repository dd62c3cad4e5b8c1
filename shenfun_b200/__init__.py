"""shenfun_b200 -- B200-native tensor-product transform-and-solve path with
shenfun's API (TensorProductSpace.forward/backward, Function/Array,
inner/project, la solvers).  Kernels live in libshenb200.so (sm_100a, C ABI in
include/shenb200.h); this package is the host-side mirror of the reference's
Python interface for that path.
"""
from .comm import comm, Comm, COMM_SELF, init_from_env
from .config import config
from .spectralbase import SpectralBase, Domain
from .fourier import R2C, C2C
from .functionspace import FunctionSpace
from .tensorproductspace import TensorProductSpace, VectorSpace, TensorSpace, CompositeSpace
from .arrays import Function, Array
from . import ops
from .forms import (Expr, TestFunction, TrialFunction, Dx, grad, div, curl, inner, project, Inner, Project)
from .matrices import TPMatrix, SparseMatrix, get_simplified_tpmatrices
from .integrators import RK4, IntegratorBase, PDEIMEXRK, IMEXRK011, IMEXRK111, IMEXRK222, IMEXRK443
from .utilities import cross, apply_mask, dx
from . import la, chebyshev, legendre
from .io import Checkpoint
from . import dlt
legendre.dlt = dlt          # shenfun.legendre.dlt: Leg2Cheb / Cheb2Leg / DLT
from .polybase import BoundaryConditions
from ._lib import ShenB200Error

__version__ = '0.1.0'
