"""Pin the numpy oracle (oracle/fourier.py, oracle/tensorproduct.py) against the
golden vectors produced by the reference's own sources (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import tensorproduct as otp

CASES = ['f3d_14_15_16', 'f3d_16_16_16', 'f3d_12_9_10', 'f2d_16_12', 'f2d_c_10_12']
TOL = 1e-14


def descs(g, pf=1.0):
    N = g['N']
    dts = str(g['dtypes'])
    return [('C2C' if dt == 'D' else 'R2C', int(n), int(np.floor(pf*n))) for n, dt in zip(N, dts)]


@pytest.mark.parametrize('name', CASES)
def test_forward_backward_match_reference(name):
    g = load_golden(name)
    d = descs(g)
    assert rel_l2(otp.forward(g['u'], d), g['u_hat']) < TOL
    assert rel_l2(otp.backward(g['v_hat'], d), g['v']) < TOL


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('pf', [1.5, 2.0])
def test_padded_match_reference(name, pf):
    g = load_golden(name)
    tag = str(pf).replace('.', '')
    d = descs(g, pf)
    assert rel_l2(otp.backward(g['v_hat'], d), g[f'v_pad{tag}']) < TOL
    assert rel_l2(otp.forward(g[f'w_pad{tag}'], d), g[f'w_hat_pad{tag}']) < TOL


@pytest.mark.parametrize('N', [(14, 15, 16), (64, 64, 64)])
def test_fourier_poisson3d_matches_reference_demo(N):
    """BASELINE config 0: the oracle's restatement of demo/fourier_poisson3D.py against the
    reference's own run of it."""
    g = load_golden('fourier_poisson3d')
    tag = 'x'.join(str(n) for n in N)
    f_hat, u_hat = otp.fourier_poisson3d(N)
    assert rel_l2(f_hat, g[f'f_hat_{tag}']) < 1e-13
    assert rel_l2(u_hat, g[f'u_hat_{tag}']) < 1e-13


@pytest.mark.parametrize('pf,tag', [(1.0, 'plain'), (1.5, 'pad15')])
def test_navier_stokes_oracle_matches_reference_demo(pf, tag):
    """oracle/navier_stokes.py (the CPU baseline of the `ns` bench key) against the reference's own
    run of demo/NavierStokes.py at 16^3 after 1 and 5 RK4 steps."""
    from oracle.navier_stokes import NavierStokes
    g = load_golden('ns_16')
    s = NavierStokes((16, 16, 16), nu=0.000625, dt=0.01, padding_factor=pf).initialize()
    s.step()
    assert rel_l2(s.U_hat, g[f'U_hat_{tag}_1']) < 1e-12
    for _ in range(4):
        s.step()
    assert rel_l2(s.U_hat, g[f'U_hat_{tag}_5']) < 1e-12
