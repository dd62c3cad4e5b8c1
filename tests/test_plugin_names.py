"""The by-name plugin boundary (shenfun/optimization/__init__.py:47-84): name resolution and the
exported surface.  No device work here; the kernels themselves are checked in test_gpu_plugin.py."""
import pytest

# the hot-path names of optimization/cython/__init__.py:1-18
EXPORTED = """TDMA_LU TDMA_Solve PDMA_LU PDMA_Solve LU_Helmholtz Solve_Helmholtz LU_Biharmonic
Biharmonic_factor_pr Biharmonic_Solve TDMA_O_Solve TDMA_O_LU Poisson_Solve_ADD FDMA_Solve TwoDMA_Solve
ThreeDMA_Solve FDMA_LU DiagMA_Solve TDMA_inner_solve TDMA_O_inner_solve DiagMA_inner_solve PDMA_inner_solve
FDMA_inner_solve TwoDMA_inner_solve ThreeDMA_inner_solve HeptaDMA_inner_solve HeptaDMA_Solve HeptaDMA_LU
SolverGeneric1ND_solve_data Helmholtz_matvec Biharmonic_matvec outer2D outer3D apply_mask cross2D cross3D
evaluate_expansion_all scalar_product""".split()


def test_exported_names():
    from shenfun_b200.optimization import b200
    missing = [n for n in EXPORTED if not callable(getattr(b200, n, None))]
    assert not missing, missing
    assert callable(b200.transforms.scalar_product) and callable(b200.transforms.evaluate_expansion_all)


class TDMA:
    """Stand-in for shenfun.la.TDMA: the plugin is looked up as `TDMA_Solve` / `TDMA_LU`."""
    @staticmethod
    def Solve(u, data, axis):
        raise AssertionError('python body must not run')

    @staticmethod
    def LU(data):
        raise AssertionError('python body must not run')


def test_optimizer_resolves_by_name_and_never_falls_back():
    from shenfun_b200.optimization import optimizer, runtimeoptimizer
    from shenfun_b200.optimization import b200
    assert optimizer(TDMA.Solve, wrap=False) is b200.TDMA_Solve
    assert optimizer(TDMA.LU, wrap=False) is b200.TDMA_LU

    def chebval(x, c):
        return None
    with pytest.raises(NotImplementedError):
        optimizer(chebval)
    with pytest.raises(NotImplementedError):
        runtimeoptimizer(chebval)(1, 2)
