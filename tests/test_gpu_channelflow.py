"""Channel flow (BASELINE config 5, demo/MKM.py) at 32x16x16 against the reference's own run
(tests/golden/mkm_32_16_16.npz): initial condition, one and three IMEXRK222 time steps, both
convection forms."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from shenfun_b200.demos.channelflow import MKM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('conv', [0, 1])
def test_mkm_matches_reference(conv):
    g = load_golden('mkm_32_16_16')
    c = MKM(N=(32, 16, 16), Re=180., dt=0.001, conv=conv, padding_factor=(1.5, 1.5, 1.5), timestepper='IMEXRK222')
    t, tstep = c.initialize()
    assert rel_l2(c.u_.get(), g[f'u0_conv{conv}']) < 1e-12
    assert rel_l2(c.g_.get(), g[f'g0_conv{conv}']) < 1e-12
    t, tstep = c.solve(t=t, tstep=tstep, end_time=0.001)
    assert rel_l2(c.H_.get(), g[f"H1_conv{conv}"]) < 1e-12
    assert rel_l2(c.u_.get(), g[f"u1_conv{conv}"]) < 1e-12
    t, tstep = c.solve(t=t, tstep=tstep, end_time=0.003)
    assert rel_l2(c.u_.get(), g[f"u3_conv{conv}"]) < 1e-12
    assert rel_l2(c.g_.get(), g[f"g3_conv{conv}"]) < 1e-12


def test_dot_grad3d_kernel():
    """H_i = sum_j u_j du_i/dx_j in one pass == the numpy expression of demo/ChannelFlow.py:213-215."""
    import torch
    from shenfun_b200 import ops
    rng = np.random.default_rng(3)
    shape = (5, 6, 7)
    u = rng.standard_normal((3,)+shape)
    g = [rng.standard_normal(shape) for _ in range(9)]
    ref = np.stack([u[0]*g[3*i]+u[1]*g[3*i+1]+u[2]*g[3*i+2] for i in range(3)])
    out = torch.empty((3,)+shape, dtype=torch.float64, device='cuda')
    ops.dot_grad3d(torch.from_numpy(u).cuda(), [torch.from_numpy(x).cuda() for x in g], out)
    assert np.abs(out.cpu().numpy()-ref).max() < 1e-14


@pytest.mark.gpu
def test_kmm_vw_kernel_matches_the_reference_expression():
    """shb_kmm_vw against demo/ChannelFlow.py:238-239 evaluated with torch."""
    import torch
    from shenfun_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(3)
    n0, n1, n2 = 19, 12, 7
    f = torch.randn(n0, n1, n2, dtype=torch.complex128, device='cuda', generator=g)
    gg = torch.randn(n0, n1, n2, dtype=torch.complex128, device='cuda', generator=g)
    a = torch.randn(1, n1, n2, dtype=torch.float64, device='cuda', generator=g)
    b = torch.randn(1, n1, n2, dtype=torch.float64, device='cuda', generator=g)
    u = torch.zeros(3, n0, n1, n2, dtype=torch.complex128, device='cuda')
    ops.kmm_vw(f, gg, a, b, u[1], u[2])
    v_ref = 1j*(a*f+b*gg)
    w_ref = 1j*(b*f-a*gg)
    assert (u[1]-v_ref).abs().max().item() <= 1e-14*v_ref.abs().max().item()
    assert (u[2]-w_ref).abs().max().item() <= 1e-14*w_ref.abs().max().item()
    assert u[0].abs().max().item() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize('axis,order', [(1, 1), (2, 1), (1, 2), (2, 3)])
def test_project_of_a_periodic_derivative_takes_the_pointwise_path(axis, order):
    """Project(Dx(u, periodic axis, k), space of u) = u_hat times (i k)^order: the point-wise shortcut against
    the generic banded mat-vec + mass solve (SHB_PROJECT_SHORTCUT=0)."""
    import os
    import torch
    import shenfun_b200 as sf
    SD = sf.FunctionSpace(32, 'C', bc=(0, 0))
    T = sf.TensorProductSpace(sf.comm, (SD, sf.FunctionSpace(12, 'F', dtype='D'), sf.FunctionSpace(10, 'F', dtype='d')))
    u = sf.Function(T)
    u.v.normal_()
    u.v[SD.dim():] = 0
    P = sf.Project(sf.Dx(u, axis, order), T)
    assert P._short is not None
    fast = P().get().copy()
    os.environ['SHB_PROJECT_SHORTCUT'] = '0'
    try:
        Pg = sf.Project(sf.Dx(u, axis, order), T)
    finally:
        del os.environ['SHB_PROJECT_SHORTCUT']
    assert Pg._short is None
    ref = Pg().get()
    assert np.abs(fast-ref).max() <= 1e-12*np.abs(ref).max()
    # a derivative along the non-periodic axis keeps the generic path
    assert sf.Project(sf.Dx(u, 0, 1), sf.TensorProductSpace(sf.comm, (sf.FunctionSpace(32, 'C'), T.bases[1], T.bases[2])))._short is None
