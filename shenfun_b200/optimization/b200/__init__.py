"""Names exported to the by-name plugin boundary (mirrors optimization/cython/__init__.py:1-18).

Arguments are CUDA `torch.Tensor`s (zero copy, in place) or the reference's own numpy arrays
(staged to the device and written back in place -- the literal drop-in, PCIe-bound).
"""
from .la import (TDMA_LU, TDMA_Solve, TDMA_inner_solve, TDMA_O_LU, TDMA_O_Solve, TDMA_O_inner_solve,
                 PDMA_LU, PDMA_Solve, PDMA_inner_solve, FDMA_LU, FDMA_Solve, FDMA_inner_solve,
                 TwoDMA_Solve, TwoDMA_inner_solve, ThreeDMA_Solve, ThreeDMA_inner_solve,
                 HeptaDMA_LU, HeptaDMA_Solve, HeptaDMA_inner_solve, DiagMA_Solve, DiagMA_inner_solve,
                 SolverGeneric1ND_solve_data, LU_Helmholtz, Solve_Helmholtz, Poisson_Solve_ADD,
                 LU_Biharmonic, Biharmonic_factor_pr, Biharmonic_Solve)
from .Matvec import Helmholtz_matvec, Biharmonic_matvec, Tridiagonal_matvec, Pentadiagonal_matvec
from .pointwise import cross2D, cross3D, outer2D, outer3D, apply_mask
from . import transforms
from .transforms import scalar_product, evaluate_expansion_all
