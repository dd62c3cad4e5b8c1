"""Host half of `TensorProductSpace.eval` (tensorproductspace.py:668-854): the per-axis basis
matrices `evaluate_basis_all` and the conjugate-half rule of the real-to-complex axis, contracted
with numpy here, reproduce the reference's backward transforms at mesh points (golden vectors)."""
import numpy as np
import pytest

from conftest import load_golden, BCS
import shenfun_b200 as sf


def host_eval(bases, pts, uh):
    letters = 'abcd'[:len(bases)]
    mats, r2c, cm = [], None, None
    for ax, b in enumerate(bases):
        E = np.asarray(b.evaluate_basis_all(pts[ax])).astype(complex)
        if b.family() == 'fourier' and b.short_name() == 'R2C':
            r2c = ax
            k = np.arange(b.spectral_len())
            last = b.N//2 if b.N % 2 == 0 else b.N//2+1
            cm = ((k >= 1) & (k < last)).astype(float)
        mats.append(E)
    expr = ','.join('p'+c for c in letters)+','+letters+'->p'
    val = np.einsum(expr, *mats, uh)
    if r2c is not None:
        m2 = list(mats)
        m2[r2c] = mats[r2c]*cm
        val = val+np.conj(np.einsum(expr, *m2, uh))
    return val.real


@pytest.mark.parametrize('name', ['f3d_14_15_16', 'f3d_12_9_10'])
def test_fourier_basis_evaluation(name):
    g = load_golden(name)
    N = [int(n) for n in g['N']]
    bases = [sf.FunctionSpace(N[0], 'F', dtype='D'), sf.FunctionSpace(N[1], 'F', dtype='D'),
             sf.FunctionSpace(N[2], 'F', dtype='d')]
    xs = [2*np.pi*np.arange(n)/n for n in N]
    idx = np.random.default_rng(0).integers(0, N, size=(20, 3))
    pts = np.array([[xs[a][i[a]] for i in idx] for a in range(3)])
    ref = np.array([g['v'][tuple(i)] for i in idx])
    assert np.abs(host_eval(bases, pts, g['v_hat'])-ref).max() < 1e-13*np.abs(ref).max()


@pytest.mark.parametrize('fam', ['C', 'L'])
@pytest.mark.parametrize('name', ['dirichlet', 'biharmonic', 'neumann'])
def test_polynomial_basis_evaluation(fam, name):
    g = load_golden('poly3d')
    tag = f'{fam}_{name}_first'
    SD = sf.FunctionSpace(14, fam, bc=BCS[name])
    bases = [SD, sf.FunctionSpace(12, 'F', dtype='D'), sf.FunctionSpace(10, 'F', dtype='d')]
    xs = [SD.points_and_weights(14)[0], 2*np.pi*np.arange(12)/12, 2*np.pi*np.arange(10)/10]
    idx = np.random.default_rng(1).integers(0, [14, 12, 10], size=(15, 3))
    pts = np.array([[xs[a][i[a]] for i in idx] for a in range(3)])
    ref = np.array([g[f'bw_{tag}'][tuple(i)] for i in idx])
    assert np.abs(host_eval(bases, pts, g[f'vh_{tag}'])-ref).max() < 1e-13*np.abs(ref).max()


def test_boundary_lift_rows_of_the_basis_matrix():
    g = load_golden('poly1d_C')
    S = sf.FunctionSpace(16, 'C', bc=BCS['inhom'])
    x = S.points_and_weights(16)[0]
    assert np.abs(S.evaluate_basis_all(x) @ g['vh_inhom_16']-g['bw_inhom_16']).max() < 1e-13
