"""CPU: pin the oracle for Chebyshev/Legendre transforms and line solvers, and the host-side
matrix/stencil assembly of shenfun_b200, against golden vectors from the reference."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2, BCS, oracle_spec
import shenfun_b200 as sf
from oracle import polybases as opb
from oracle import la as ola

HOMOG = ['ortho', 'dirichlet', 'neumann', 'biharmonic', 'dn']


@pytest.mark.parametrize('fam', ['C', 'L'])
@pytest.mark.parametrize('name', HOMOG)
def test_stencil_and_matrices_match_reference(fam, name):
    g = load_golden('poly1d_'+fam)
    S = sf.FunctionSpace(16, fam, bc=BCS[name])
    tag = f'{name}_16'
    if f'K_{tag}' in g:
        assert np.abs(S.stencil_matrix().diags()-g[f'K_{tag}'][:S.dim()]).max() < 1e-15
    for (a, b) in ((0, 0), (0, 1), (0, 2), (1, 1), (0, 4), (2, 2)):
        key = f'M{a}{b}_{tag}'
        if key not in g:
            continue
        ref = g[key]
        M = S.spectral_matrix(a, b).diags()
        assert M.shape == ref.shape
        assert np.abs(M-ref).max() <= 1e-13*max(1.0, np.abs(ref).max()), (a, b)


@pytest.mark.parametrize('fam', ['C', 'L'])
@pytest.mark.parametrize('name', HOMOG)
@pytest.mark.parametrize('N', [16, 15])
def test_oracle_1d_transforms(fam, name, N):
    g = load_golden('poly1d_'+fam)
    S = sf.FunctionSpace(N, fam, bc=BCS[name])
    spec = oracle_spec(S)
    tag = f'{name}_{N}'
    # rows >= dim of the reference's un-padded scalar product hold stale work-array values
    # (spectralbase.py:1363-1367 skips the zero fill when nothing is truncated): compare dofs
    d = S.dim()
    assert rel_l2(opb.scalar_product_axis(g[f'u_{tag}'], 0, spec)[:d], g[f'sp_{tag}'][:d]) < 1e-13
    assert rel_l2(opb.forward_axis(g[f'u_{tag}'], 0, spec), g[f'fw_{tag}']) < 1e-12
    assert rel_l2(opb.backward_axis(g[f'vh_{tag}'], 0, spec), g[f'bw_{tag}']) < 1e-13
    Sp = S.get_dealiased(1.5)
    specp = oracle_spec(Sp)
    assert rel_l2(opb.backward_axis(g[f'vh_{tag}'], 0, specp), g[f'bwpad_{tag}']) < 1e-13
    assert rel_l2(opb.forward_axis(g[f'w_{tag}'], 0, specp), g[f'fwpad_{tag}']) < 1e-12


@pytest.mark.parametrize('name', ['TDMA', 'PDMA', 'TDMA_L', 'PDMA_L'])
@pytest.mark.parametrize('axis', [0, 1, 2])
def test_oracle_banded_solvers(name, axis):
    g = load_golden('solvers')
    u = ola.solve_axis(g[f'mat_{name}'], g[f'f_{name}_{axis}'], axis)
    assert rel_l2(u, g[f'u_{name}_{axis}']) < 1e-12


def _poisson_mats(fam):
    SD = sf.FunctionSpace(14, fam, bc=(0, 0))
    A = SD.spectral_matrix(0, 2).diags()
    B = SD.mass_matrix().diags()
    k1 = np.fft.fftfreq(12, 1./12)[None, :, None]
    k2 = np.fft.rfftfreq(10, 1./10)[None, None, :]
    return A, B, k1, k2


def test_oracle_generic_1nd_and_helmholtz():
    g = load_golden('solvers')
    A, B, k1, k2 = _poisson_mats('L')
    f = g['poisson_leg_f']
    u = ola.solve_generic_1nd([A, B], [np.ones((1, 1, 1)), -(k1**2+k2**2)], f, 0, pin_zero_line=(0,))
    assert rel_l2(u, g['poisson_leg_u']) < 1e-12
    u = ola.solve_generic_1nd([A, B], [np.ones((1, 1, 1)), -(k1**2+k2**2)-2.5], f, 0)
    assert rel_l2(u, g['helm_leg_u']) < 1e-12
    A, B, k1, k2 = _poisson_mats('C')
    u = ola.helmholtz_solve(A, B, np.ones((1, 1, 1)), -(k1**2+k2**2)-2.5, g['poisson_cheb_f'], 0)
    assert rel_l2(u, g['helm_cheb_u']) < 1e-12


def test_oracle_chebyshev_biharmonic():
    g = load_golden('solvers')
    SB = sf.FunctionSpace(16, 'C', bc=(0, 0, 0, 0))
    S = SB.spectral_matrix(0, 4).diags()
    A = SB.spectral_matrix(0, 2).diags()
    B = SB.mass_matrix().diags()
    k1 = np.fft.fftfreq(12, 1./12)[None, :, None]
    k2 = np.fft.rfftfreq(10, 1./10)[None, None, :]
    kk = k1**2+k2**2
    u = ola.biharmonic_solve(S, A, B, np.ones((1, 1, 1)), -2*kk, kk**2, g['cheb_biharm_f'], 0)
    assert rel_l2(u, g['cheb_biharm_u']) < 1e-12


@pytest.mark.parametrize('name', ['ortho', 'dirichlet', 'neumann', 'biharmonic', 'dn'])
@pytest.mark.parametrize('N', [16, 15])
def test_chebyshev_lobatto_operator_matrices_host(name, N):
    """Host logic of quad='GL': the operator matrices handed to the device GEMM and the mass
    matrix reproduce the reference's DCT-I based transforms (tests/golden/poly1d_CGL.npz)."""
    import shenfun_b200 as sf
    g = load_golden('poly1d_CGL')
    S = sf.FunctionSpace(N, 'C', bc=BCS[name], quad='GL')
    tag = f'{name}_{N}'
    d = S.dim()
    F, B = S._gl_operators_host()
    sp = F @ g[f'u_{tag}']
    assert rel_l2(sp[:d], g[f'sp_{tag}'][:d]) < 1e-13
    assert rel_l2(B @ g[f'vh_{tag}'], g[f'bw_{tag}']) < 1e-13
    Fp, Bp = S.get_dealiased(1.5)._gl_operators_host()
    assert rel_l2(Bp @ g[f'vh_{tag}'], g[f'bwpad_{tag}']) < 1e-13
    M = S.mass_matrix().diags()
    fw = np.zeros(N)
    fw[:d] = np.linalg.solve(M[:d, :d], sp[:d])
    assert rel_l2(fw, g[f'fw_{tag}']) < 1e-12
    if N == 16:
        assert np.abs(M[:d, :d]-g[f'M00_{tag}'][:d, :d]).max() < 1e-13


@pytest.mark.parametrize('name', ['CTSD', 'CTSB', 'CTT', 'CTSN', 'ASDSD', 'ATT', 'SSBSB'])
def test_separable_tail_detection_host(name):
    """Host logic behind `shb_tail_matvec`: the structure found in the Chebyshev derivative
    matrices reproduces the dense product when applied as band + running sums per parity; matrices
    without a rank-one tail are left to the other paths."""
    import shenfun_b200 as sf
    from shenfun_b200.la import separable_tail
    N = 33
    TC = sf.FunctionSpace(N, 'C')
    SD = sf.FunctionSpace(N, 'C', bc=(0, 0))
    SB = sf.FunctionSpace(N, 'C', bc=(0, 0, 0, 0))
    SN = sf.FunctionSpace(N, 'C', bc=BCS['neumann'])
    test, trial, a, b = {'CTSD': (TC, SD, 0, 1), 'CTSB': (TC, SB, 0, 1), 'CTT': (TC, TC, 0, 1), 'CTSN': (TC, SN, 0, 1),
                         'ASDSD': (SD, SD, 0, 2), 'ATT': (TC, TC, 0, 2), 'SSBSB': (SB, SB, 0, 4)}[name]
    D = test.spectral_matrix(a, b, trial).diags()
    found = separable_tail(D)
    if name in ('ATT', 'SSBSB'):
        assert found is None
        return
    w, r, t, band = found
    n_rows, n_cols = D.shape
    u = np.random.default_rng(0).standard_normal(n_cols)
    out = np.zeros(n_rows)
    for o, v in band.items():
        for i, val in enumerate(v):
            k = i if o >= 0 else i-o
            out[k] += val*u[k+o]
    P = [0.0, 0.0]
    for k in range(n_rows-1, -1, -1):
        j = k+w
        if 0 <= j < n_cols:
            P[(k+w) & 1] += t[j]*u[j]
        out[k] += r[k]*P[(k+w) & 1]
    assert rel_l2(out, D @ u) < 1e-13
