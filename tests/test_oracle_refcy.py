"""Pin the numpy oracle against the reference's OWN compiled kernels: oracle/_ref/refcy holds
shenfun/optimization/cython/{la,Matvec,transforms}.pyx cythonised unchanged (oracle/build_ref.sh).
These are the kernels `config['optimization']['mode'] = 'cython'` selects in the reference
(shenfun/optimization/__init__.py:47-84).  Skipped where oracle/_ref/refcy has not been built."""
import numpy as np
import pytest

from conftest import rel_l2
import shenfun_b200 as sf
from oracle import la as ola, polybases as opb, refcy

pytestmark = pytest.mark.skipif(not refcy.available(), reason='oracle/_ref/refcy not built')


def _dia(M, offsets):
    """scipy-DIA `data` of a dense banded matrix (upper diagonals right-aligned, la.py:347-349)."""
    n = M.shape[0]
    data = np.zeros((len(offsets), n))
    for r, off in enumerate(offsets):
        d = np.diagonal(M, off)
        if off >= 0:
            data[r, off:] = d
        else:
            data[r, :n+off] = d
    return data


@pytest.mark.parametrize('kind,bc,offs', [('TDMA', (0, 0), (-2, 0, 2)), ('PDMA', (0, 0, 0, 0), (-4, -2, 0, 2, 4))])
@pytest.mark.parametrize('fam', ['C', 'L'])
@pytest.mark.parametrize('axis', [0, 1, 2])
def test_banded_solvers_against_reference_cython(kind, bc, offs, fam, axis):
    la = refcy.load('la')
    rng = np.random.default_rng(3)
    n = 34
    S = sf.FunctionSpace(n, fam, bc=bc)
    M = S.mass_matrix().diags()
    d = S.dim()
    data = _dia(M[:d, :d], offs)
    getattr(la, f'{kind}_LU')(data)
    shape = [5, 6, 7]
    shape[axis] = d
    b = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
    u = b.copy()
    getattr(la, f'{kind}_Solve')(u, data, axis)
    assert rel_l2(ola.solve_axis(M[:d, :d], b, axis), u) < 1e-12


@pytest.mark.parametrize('axis', [0, 1, 2])
@pytest.mark.parametrize('n', [18, 130])
def test_helmholtz_against_reference_cython(axis, n):
    la = refcy.load('la')
    rng = np.random.default_rng(5)
    SD = sf.FunctionSpace(n, 'C', bc=(0, 0))
    d = SD.dim()
    Ad = SD.spectral_matrix(0, 2).diags()[:d, :d]
    Bd = SD.mass_matrix().diags()[:d, :d]
    A = {k: np.ascontiguousarray(np.diagonal(Ad, k)) for k in (0, 2, 4)}
    B = {k: np.ascontiguousarray(np.diagonal(Bd, k)) for k in (-2, 0, 2)}
    shape = [4, 5, 6]
    shape[axis] = n
    sshape = list(shape)
    sshape[axis] = 1
    alfa = np.ones(sshape)
    beta = -rng.uniform(0.0, 40.0, sshape)
    d0, d1, d2, L = (np.zeros(shape) for _ in range(4))
    la.LU_Helmholtz(A, B, alfa, beta, False, d0, d1, d2, L, axis)
    b = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
    sl = [slice(None)]*3
    sl[axis] = slice(d, None)
    b[tuple(sl)] = 0
    u = np.zeros_like(b)
    la.Solve_Helmholtz(b, u, False, d0, d1, d2, L, axis)
    ref = ola.helmholtz_solve(Ad, Bd, alfa, beta, b, axis)
    sd = [slice(None)]*3
    sd[axis] = slice(0, d)
    assert rel_l2(ref[tuple(sd)], u[tuple(sd)]) < 1e-12


@pytest.mark.parametrize('tag', ['leg', 'jac', 'chebu'])
@pytest.mark.parametrize('axis', [0, 1, 2])
def test_recursive_transforms_cython_vs_numba_reference(tag, axis):
    """The reference's Cython `transforms.scalar_product` / `evaluate_expansion_all` on the inputs of
    tests/golden/plugin.npz (whose outputs came from the reference's numba kernels): the two
    reference implementations agree, and the golden file is what the device plugin is tested with."""
    from conftest import load_golden
    tr = refcy.load('transforms')
    g = load_golden('plugin')
    x, w, a = g[f'tr_{tag}_x'], g[f'tr_{tag}_w'], g[f'tr_{tag}_a']
    uin = np.ascontiguousarray(g[f'tr_{tag}_{axis}_real_in'])
    sp = np.zeros_like(uin)
    tr.scalar_product(uin, sp, x, w, axis, a)
    assert rel_l2(sp, g[f'tr_{tag}_{axis}_real_sp']) < 1e-12
    ev = np.zeros_like(uin)
    tr.evaluate_expansion_all(np.ascontiguousarray(g[f'tr_{tag}_{axis}_real_sp']), ev, x, axis, a)
    assert rel_l2(ev, g[f'tr_{tag}_{axis}_real_ev']) < 1e-12
