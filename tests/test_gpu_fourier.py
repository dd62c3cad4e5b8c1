"""Device parity for the Fourier tensor-product path (through the public API,
which calls the C ABI): against golden vectors from the reference, against the
numpy oracle on seeded inputs, and size-independent properties at 512^3."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
import shenfun_b200 as sf
from oracle import tensorproduct as otp

pytestmark = pytest.mark.gpu
TOL = 1e-12   # north_star: relative L2 <= 1e-12 in fp64
CASES = ['f3d_14_15_16', 'f3d_16_16_16', 'f3d_12_9_10', 'f2d_16_12', 'f2d_c_10_12']


def space(N, dts, pf=1.0):
    bases = [sf.FunctionSpace(int(n), 'F', dtype=dt, padding_factor=pf) for n, dt in zip(N, dts)]
    return sf.TensorProductSpace(sf.comm, bases)


def descs(N, dts, pf=1.0):
    return [('C2C' if dt == 'D' else 'R2C', int(n), int(np.floor(pf*n))) for n, dt in zip(N, dts)]


@pytest.mark.parametrize('name', CASES)
def test_golden_forward_backward(name):
    g = load_golden(name)
    T = space(g['N'], str(g['dtypes']))
    uh = T.forward(g['u']).get()
    assert rel_l2(uh, g['u_hat']) < TOL
    v = T.backward(sf.Function(T, buffer=g['v_hat'])).get()
    assert rel_l2(v, g['v']) < TOL


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('pf', [1.5, 2.0])
def test_golden_padded(name, pf):
    g = load_golden(name)
    tag = str(pf).replace('.', '')
    T = space(g['N'], str(g['dtypes']))
    Tp = T.get_dealiased((pf,)*len(g['N']))
    vh = sf.Function(T, buffer=g['v_hat'])
    v = vh.backward(padding_factor=(pf,)*len(g['N'])).get()
    assert rel_l2(v, g[f'v_pad{tag}']) < TOL
    wh = Tp.forward(g[f'w_pad{tag}']).get()
    assert rel_l2(wh, g[f'w_hat_pad{tag}']) < TOL


@pytest.mark.parametrize('N,dts', [((64, 64, 64), 'DDd'), ((128, 64, 32), 'DDd'), ((48, 40, 36), 'DDd'),
                                   ((96, 128), 'Dd'), ((256,), 'd'), ((243,), 'D'), ((34, 21, 26), 'DDD')])
@pytest.mark.parametrize('pf', [1.0, 1.5])
def test_oracle_seeded(N, dts, pf):
    rng = np.random.default_rng(7)
    T = space(N, dts)
    Tp = T.get_dealiased((pf,)*len(N))
    d = descs(N, dts, pf)
    u = rng.standard_normal(Tp.shape(False))
    if dts[-1] == 'D':
        u = u + 1j*rng.standard_normal(Tp.shape(False))
    uh = Tp.forward(u)
    assert rel_l2(uh.get(), otp.forward(u, d)) < TOL
    vh = rng.standard_normal(T.shape(True)) + 1j*rng.standard_normal(T.shape(True))
    v = Tp.backward(sf.Function(T, buffer=vh))
    assert rel_l2(v.get(), otp.backward(vh, d)) < TOL


def test_vector_space_batched():
    rng = np.random.default_rng(3)
    N, dts = (32, 24, 20), 'DDd'
    T = space(N, dts)
    TV = sf.VectorSpace(T)
    U = rng.standard_normal(TV.shape(False))
    Uh = TV.forward(sf.Array(TV, buffer=U))
    ref = otp.forward(U, descs(N, dts), lead=1)
    assert rel_l2(Uh.get(), ref) < TOL
    Ub = TV.backward(Uh)
    assert rel_l2(Ub.get(), U) < TOL
    Up = Uh.backward(padding_factor=1.5)
    assert tuple(Up.shape) == (3, 48, 36, 30)
    assert rel_l2(Up.get(), otp.backward(ref, descs(N, dts, 1.5), lead=1)) < TOL


def test_forward_preserves_input_and_roundtrip():
    T = space((32, 32, 32), 'DDd')
    u = sf.Array(T)
    u[:] = torch.randn(32, 32, 32, dtype=torch.float64, device='cuda')
    u0 = u.copy()
    uh = u.forward()
    assert torch.equal(u.v, u0.v)
    uh0 = uh.copy()
    ub = uh.backward()
    assert torch.equal(uh.v, uh0.v)
    assert rel_l2(ub.get(), u0.get()) < TOL


def test_dealias_direct():
    rng = np.random.default_rng(5)
    from oracle import fourier as of
    N = (16, 18, 12)
    bases = [sf.FunctionSpace(n, 'F', dtype=dt, dealias_direct=True) for n, dt in zip(N, 'DDd')]
    T = sf.TensorProductSpace(sf.comm, bases)
    vh = rng.standard_normal(T.shape(True)) + 1j*rng.standard_normal(T.shape(True))
    v = T.backward(sf.Function(T, buffer=vh)).get()
    ref = of.c2c_backward(vh, 0, dealias_direct=True)
    ref = of.c2c_backward(ref, 1, dealias_direct=True)
    ref = of.c2r_backward(ref, 2, N[2], dealias_direct=True)
    assert rel_l2(v, ref) < TOL


@pytest.mark.parametrize('pf', [1.0, 1.5])
def test_full_size_512_properties(pf):
    """BASELINE size: round trip, linearity and Parseval at 512^3 (oracle-free)."""
    N = (512, 512, 512)
    T = space(N, 'DDd')
    Tp = T.get_dealiased((pf,)*3)
    g = torch.Generator(device='cuda').manual_seed(1234)
    vh = sf.Function(T)
    u = torch.randn(N, dtype=torch.float64, device='cuda', generator=g)
    T.forward(u, vh)                      # a Hermitian-consistent spectrum
    vh.mask_nyquist()
    v = Tp.backward(vh)
    vh2 = Tp.forward(v)
    err = (torch.linalg.vector_norm(vh2.v-vh.v)/torch.linalg.vector_norm(vh.v)).item()
    assert err < TOL
    # Parseval (tests/test_fourier.py:5-21): mean(u^2) = sum |u_hat|^2 with Hermitian weights
    w = torch.full((vh.shape[2],), 2.0, dtype=torch.float64, device='cuda')
    w[0] = 1.0
    w[-1] = 1.0
    e_spec = (vh.v.abs()**2*w).sum().item()
    e_phys = (v.v**2).mean().item()
    assert abs(e_spec-e_phys)/e_phys < 1e-12
    # linearity
    a = 0.37
    v2 = Tp.backward(sf.Function(T, buffer=(vh.v*a)))
    assert (torch.linalg.vector_norm(v2.v-a*v.v)/torch.linalg.vector_norm(v.v)).item() < 1e-14


@pytest.mark.parametrize('N', [(14, 15, 16), (64, 64, 64)])
def test_fourier_poisson3d_demo(N):
    """BASELINE config 0 (demo/fourier_poisson3D.py): solution coefficients against the reference's
    own run of the demo, and the demo's own assertion on the L2 error (:49)."""
    from shenfun_b200.demos import fourier_poisson3d
    g = load_golden('fourier_poisson3d')
    tag = 'x'.join(str(n) for n in N)
    u_hat, err = fourier_poisson3d.solve(N)
    assert err < 1e-6
    assert abs(err-float(g[f'error_{tag}'])) < 1e-12
    assert rel_l2(u_hat.get(), g[f'u_hat_{tag}']) < TOL


def test_eval_at_points_matches_backward_on_the_mesh():
    """`TensorProductSpace.eval` (tensorproductspace.py:668-854) at mesh points == backward."""
    import shenfun_b200 as sf
    rng = np.random.default_rng(8)
    N = (10, 9, 12)
    T = sf.TensorProductSpace(sf.comm, [sf.FunctionSpace(N[0], 'F', dtype='D'), sf.FunctionSpace(N[1], 'F', dtype='D'),
                                        sf.FunctionSpace(N[2], 'F', dtype='d')])
    u = rng.standard_normal(N)
    uh = T.forward(u)
    xs = [2*np.pi*np.arange(n)/n for n in N]
    idx = rng.integers(0, N, size=(25, 3))
    pts = np.array([[xs[a][i[a]] for i in idx] for a in range(3)])
    ref = np.array([u[tuple(i)] for i in idx])
    assert np.abs(T.eval(pts, uh)-ref).max() < 1e-12
    SD = sf.FunctionSpace(14, 'C', bc=(0, 0))
    Tc = sf.TensorProductSpace(sf.comm, (SD, sf.FunctionSpace(8, 'F', dtype='D'), sf.FunctionSpace(6, 'F', dtype='d')))
    vh = np.zeros(Tc.shape(True), dtype=complex)
    vh[:SD.dim()] = rng.standard_normal((SD.dim(), 8, 4))+1j*rng.standard_normal((SD.dim(), 8, 4))
    F = sf.Function(Tc, buffer=vh)
    v = Tc.backward(F).get()
    xs = [SD.points_and_weights(14)[0], 2*np.pi*np.arange(8)/8, 2*np.pi*np.arange(6)/6]
    idx = rng.integers(0, [14, 8, 6], size=(25, 3))
    pts = np.array([[xs[a][i[a]] for i in idx] for a in range(3)])
    ref = np.array([v[tuple(i)] for i in idx])
    assert np.abs(Tc.eval(pts, F)-ref).max() < 1e-12*max(1.0, np.abs(ref).max())
