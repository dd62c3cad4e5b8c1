"""Navier-Stokes (BASELINE config 2) on the device against the reference's own run of
demo/NavierStokes.py (tests/golden/ns_16.npz) and the demo's kinetic-energy anchor."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
import shenfun_b200 as sf
from shenfun_b200.demos import navier_stokes as ns

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize('pf,tag', [(1.0, 'plain'), (1.5, 'pad15')])
@pytest.mark.parametrize('nsteps', [1, 5])
def test_reference_style_demo_matches_reference(pf, tag, nsteps):
    g = load_golden('ns_16')
    U_hat, U, k = ns.solve_reference_style((16, 16, 16), dt=0.01, end_time=nsteps*0.01, padding_factor=pf)
    assert rel_l2(U_hat.get(), g[f'U_hat_{tag}_{nsteps}']) < TOL
    assert abs(k-float(g[f'k_{tag}_{nsteps}'])) < 1e-13


@pytest.mark.parametrize('pf,tag', [(1.0, 'plain'), (1.5, 'pad15')])
@pytest.mark.parametrize('nsteps', [1, 5])
def test_fused_solver_matches_reference(pf, tag, nsteps):
    g = load_golden('ns_16')
    s = ns.FusedNS((16, 16, 16), dt=0.01, padding_factor=pf).initialize()
    for _ in range(nsteps):
        s.step()
    assert rel_l2(s.U_hat.get(), g[f'U_hat_{tag}_{nsteps}']) < TOL
    assert abs(s.energy()-float(g[f'k_{tag}_{nsteps}'])) < 1e-13


def test_taylor_green_energy_anchor_32():
    """demo/NavierStokes.py:59-61: k = 0.124953117517 at t = 0.1 (32^3, RK4) to 7 decimals."""
    s = ns.FusedNS((32, 32, 32), dt=0.01, padding_factor=1.0).initialize()
    for _ in range(10):
        s.step()
    k = s.energy()
    assert round(k-0.124953117517, 7) == 0
    assert abs(k-float(load_golden('ns_16')['k32_t01'])) < 1e-13


def test_cross_and_curl_kernels_against_numpy():
    rng = np.random.default_rng(11)
    a = rng.standard_normal((3, 9, 7, 5))
    b = rng.standard_normal((3, 9, 7, 5))
    c = torch.empty(3, 9, 7, 5, dtype=torch.float64, device='cuda')
    sf.ops.cross3d(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), c)
    assert rel_l2(c.cpu().numpy(), np.cross(a, b, axis=0)) < 1e-15
    ac = a+1j*rng.standard_normal(a.shape)
    bc = b+1j*rng.standard_normal(a.shape)
    cc = torch.empty(3, 9, 7, 5, dtype=torch.complex128, device='cuda')
    sf.ops.cross3d(torch.from_numpy(ac).cuda(), torch.from_numpy(bc).cuda(), cc)
    assert rel_l2(cc.cpu().numpy(), np.cross(ac, bc, axis=0)) < 1e-15
    k = [rng.standard_normal(n) for n in (9, 7, 5)]
    K = np.array(np.meshgrid(*k, indexing='ij'))
    out = torch.empty_like(cc)
    sf.ops.curl_hat(torch.from_numpy(ac).cuda(), out, *[torch.from_numpy(x).cuda() for x in k])
    assert rel_l2(out.cpu().numpy(), 1j*np.cross(K, ac, axis=0)) < 1e-15


def test_full_size_divergence_free_and_energy_decay():
    """128^3 with 3/2 padding: one step keeps the field solenoidal and dissipates energy
    at the viscous rate (size-independent properties; 512^3 runs in bench.py)."""
    s = ns.FusedNS((128, 128, 128), dt=0.001, padding_factor=1.5).initialize()
    k0 = s.energy()
    s.step()
    k1 = s.energy()
    K = s.T.local_wavenumbers(scaled=True, device=True)
    divu = sum(K[i]*s.U_hat.v[i] for i in range(3))
    assert divu.abs().max().item() < 1e-13
    # dk/dt = -2 nu Omega, Omega = 3/8 for Taylor-Green at t=0 -> relative change ~ 6 nu dt
    assert 0 < (k0-k1)/k0 < 1e-5


@pytest.mark.parametrize('N,pf', [((8, 8, 64), 1.0), ((6, 10, 128), 1.0), ((4, 6, 512), 1.5), ((4, 4, 256), 1.5),
                                  ((8, 8, 20), 1.0)])
def test_backward_fused_with_cross_product(N, pf):
    """`backward_cross` (C2R + cross in one kernel; the last size falls back to two passes)
    equals backward followed by numpy's cross product."""
    rng = np.random.default_rng(5)
    T, TV = ns.make_spaces(N)
    TVp = TV.get_dealiased((pf,)*3) if pf > 1 else TV
    Tp = TVp.flatten()[0]
    vh = rng.standard_normal(TV.shape(True))+1j*rng.standard_normal(TV.shape(True))
    b = rng.standard_normal(TVp.shape(False))
    Uh = sf.Function(TV, buffer=vh)
    ref = np.cross(TVp.backward(Uh).get(), b, axis=0)
    out = sf.Array(TVp)
    Tp.backward_cross(Uh, sf.Array(TVp, buffer=b), out)
    assert rel_l2(out.get(), ref) < TOL


@pytest.mark.parametrize('N,pf', [((8, 8, 64), 1.0), ((6, 10, 128), 1.0), ((4, 6, 512), 1.5), ((4, 4, 256), 1.5),
                                  ((4, 4, 1024), 1.5), ((6, 4, 1024), 1.0), ((4, 6, 256), 1.0), ((6, 4, 512), 1.0),
                                  ((8, 8, 20), 1.0)])
def test_backward_of_two_fields_fused_with_cross_product(N, pf):
    """`backward_cross_hat`: both spectral fields go through one batched backward and one
    C2R-cross kernel (the last size falls back) == two backwards and numpy's cross product."""
    import torch
    rng = np.random.default_rng(6)
    T, TV = ns.make_spaces(N)
    TVp = TV.get_dealiased((pf,)*3) if pf > 1 else TV
    Tp = TVp.flatten()[0]
    ah = rng.standard_normal(TV.shape(True))+1j*rng.standard_normal(TV.shape(True))
    bh = rng.standard_normal(TV.shape(True))+1j*rng.standard_normal(TV.shape(True))
    ref = np.cross(TVp.backward(sf.Function(TV, buffer=ah)).get(), TVp.backward(sf.Function(TV, buffer=bh)).get(),
                   axis=0)
    ab = torch.from_numpy(np.concatenate([ah, bh], axis=0)).cuda()
    out = sf.Array(TVp)
    Tp.backward_cross_hat(ab, out)
    assert rel_l2(out.get(), ref) < TOL


def test_curl_formed_inside_the_first_padded_pass():
    """`shb_fft_c2c_backward_curl` (i k x u_hat formed on the fly in the first 3/2-rule pass, the vorticity
    spectrum never stored): the kernel against numpy on random data, and two RK4 steps at 128^3 against
    the step that stores the vorticity spectrum and against the numpy oracle of demo/NavierStokes.py."""
    import os
    from shenfun_b200 import ops
    from oracle import fourier as of, navier_stokes as ons
    rng = np.random.default_rng(11)
    N, n1, n2 = 128, 5, 7
    uh = rng.standard_normal((3, N, n1, n2))+1j*rng.standard_normal((3, N, n1, n2))
    k0 = np.fft.fftfreq(N, 1./N)
    k1 = rng.standard_normal(n1)
    k2 = rng.standard_normal(n2)
    K = np.meshgrid(k0, k1, k2, indexing='ij')
    curl = 1j*np.array([K[1]*uh[2]-K[2]*uh[1], K[2]*uh[0]-K[0]*uh[2], K[0]*uh[1]-K[1]*uh[0]])
    ref = of.c2c_backward(np.concatenate([uh, curl]), 1, 192)
    out = torch.empty((6, 192, n1, n2), dtype=torch.complex128, device='cuda')
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()       # noqa: E731
    ops.fft_c2c_backward_curl(dv(uh), out, n1*n2, n2, 192, N, dv(k0), dv(k1), dv(k2))
    assert rel_l2(out.cpu().numpy(), ref) < TOL

    os.environ['SHB_NS_CURL_FUSION'] = '1'
    try:
        a = ns.FusedNS((128, 128, 128), dt=0.005, padding_factor=1.5).initialize()
        assert a._curl_fused
        for _ in range(2):
            a.step()
    finally:
        del os.environ['SHB_NS_CURL_FUSION']
    b = ns.FusedNS((128, 128, 128), dt=0.005, padding_factor=1.5).initialize()
    assert not b._curl_fused
    for _ in range(2):
        b.step()
    ua, ub = a.U_hat.get(), b.U_hat.get()
    assert rel_l2(ua, ub) < 1e-13
    o = ons.NavierStokes((128, 128, 128), nu=0.000625, dt=0.005, padding_factor=1.5).initialize()
    for _ in range(2):
        o.step()
    assert rel_l2(ua, o.U_hat) < TOL
