"""Device parity for the Chebyshev / Legendre bases, the banded line solvers, SolverGeneric1ND and
the Chebyshev Helmholtz solver against golden vectors produced by the reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, BCS, oracle_spec
import shenfun_b200 as sf
from shenfun_b200 import la
from oracle import polybases as opb
from oracle import la as ola

pytestmark = pytest.mark.gpu
TOL = 1e-12
NAMES = ['ortho', 'dirichlet', 'neumann', 'biharmonic', 'dn']


def _space1d(N, fam, bc):
    """'CGL' = Chebyshev on the Gauss-Lobatto points (quad='GL')."""
    if fam == 'CGL':
        return sf.FunctionSpace(N, 'C', bc=bc, quad='GL')
    if fam == 'LGL':      # Legendre on the Gauss-Lobatto points (legendre/bases.py:128-129, 157-164)
        return sf.FunctionSpace(N, 'L', bc=bc, quad='GL')
    return sf.FunctionSpace(N, fam, bc=bc)


@pytest.mark.parametrize('fam', ['C', 'L', 'CGL', 'LGL'])
@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('N', [16, 15])
def test_1d_transforms_golden(fam, name, N):
    g = load_golden('poly1d_'+fam)
    S = _space1d(N, fam, BCS[name])
    tag = f'{name}_{N}'
    d = S.dim()
    sp = S.scalar_product(g[f'u_{tag}']).cpu().numpy()
    assert rel_l2(sp[:d], g[f'sp_{tag}'][:d]) < TOL
    assert np.all(sp[d:] == 0)
    assert rel_l2(S.forward(g[f'u_{tag}']).cpu().numpy(), g[f'fw_{tag}']) < TOL
    assert rel_l2(S.backward(g[f'vh_{tag}']).cpu().numpy(), g[f'bw_{tag}']) < TOL
    Sp = S.get_dealiased(1.5)
    assert rel_l2(Sp.backward(g[f'vh_{tag}']).cpu().numpy(), g[f'bwpad_{tag}']) < TOL
    assert rel_l2(Sp.forward(g[f'w_{tag}']).cpu().numpy(), g[f'fwpad_{tag}']) < TOL


@pytest.mark.parametrize('fam', ['C', 'L', 'CGL', 'LGL'])
@pytest.mark.parametrize('N', [16, 15])
def test_1d_inhomogeneous_dirichlet(fam, N):
    g = load_golden('poly1d_'+fam)
    S = _space1d(N, fam, BCS['inhom'])
    tag = f'inhom_{N}'
    assert rel_l2(S.forward(g[f'u_{tag}']).cpu().numpy(), g[f'fw_{tag}']) < TOL
    assert rel_l2(S.backward(g[f'vh_{tag}']).cpu().numpy(), g[f'bw_{tag}']) < TOL


def _space(fam, name, layout, pf=1.0):
    SD = sf.FunctionSpace(14, fam, bc=BCS[name], padding_factor=pf)
    K1 = sf.FunctionSpace(12, 'F', dtype='D', padding_factor=pf)
    K2 = sf.FunctionSpace(10, 'F', dtype='d', padding_factor=pf)
    if layout == 'first':
        return sf.TensorProductSpace(sf.comm, (SD, K1, K2)), 0
    return sf.TensorProductSpace(sf.comm, (K1, SD, K2), axes=(1, 0, 2)), 1


@pytest.mark.parametrize('fam', ['C', 'L'])
@pytest.mark.parametrize('name', ['dirichlet', 'biharmonic', 'neumann'])
@pytest.mark.parametrize('layout', ['first', 'middle'])
def test_3d_spaces_golden(fam, name, layout):
    g = load_golden('poly3d')
    tag = f'{fam}_{name}_{layout}'
    T, ax = _space(fam, name, layout)
    d = T.bases[ax].dim()
    sl = [slice(None)]*3
    sl[ax] = slice(0, d)
    sl = tuple(sl)
    assert rel_l2(T.forward(g[f'u_{tag}']).get(), g[f'fw_{tag}']) < TOL
    assert rel_l2(T.scalar_product(g[f'u_{tag}']).get()[sl], g[f'sp_{tag}'][sl]) < TOL
    vh = sf.Function(T, buffer=g[f'vh_{tag}'])
    assert rel_l2(T.backward(vh).get(), g[f'bw_{tag}']) < TOL
    Tp = T.get_dealiased((1.5, 1.5, 1.5))
    assert rel_l2(Tp.backward(vh).get(), g[f'bwpad_{tag}']) < TOL
    assert rel_l2(Tp.forward(g[f'w_{tag}']).get(), g[f'fwpad_{tag}']) < TOL


@pytest.mark.parametrize('fam,N', [('C', 64), ('C', 96), ('C', 81), ('L', 64), ('L', 50)])
@pytest.mark.parametrize('name', ['dirichlet', 'biharmonic'])
def test_larger_transforms_against_oracle(fam, N, name):
    rng = np.random.default_rng(N)
    S = sf.FunctionSpace(N, fam, bc=BCS[name])
    T = sf.TensorProductSpace(sf.comm, (S, sf.FunctionSpace(12, 'F', dtype='D'), sf.FunctionSpace(8, 'F', dtype='d')))
    spec = oracle_spec(T.bases[0])
    u = rng.standard_normal((N, 12, 5))+1j*rng.standard_normal((N, 12, 5))
    ud = torch.from_numpy(u).cuda()
    out = torch.empty_like(ud)
    T.bases[0]._forward_axis(ud, out, 1, 60)
    assert rel_l2(out.cpu().numpy(), opb.forward_axis(u, 0, spec)) < TOL
    vh = np.array(u)
    vh[S.dim():] = 0
    T.bases[0]._backward_axis(torch.from_numpy(vh).cuda(), out, 1, 60)
    assert rel_l2(out.cpu().numpy(), opb.backward_axis(vh, 0, spec)) < TOL


@pytest.mark.parametrize('name,fam,bc', [('TDMA', 'C', (0, 0)), ('PDMA', 'C', (0, 0, 0, 0)),
                                         ('TDMA_L', 'L', (0, 0)), ('PDMA_L', 'L', (0, 0, 0, 0))])
@pytest.mark.parametrize('axis', [0, 1, 2])
def test_named_banded_solvers_golden(name, fam, bc, axis):
    g = load_golden('solvers')
    S = sf.FunctionSpace(16, fam, bc=bc)
    M = S.mass_matrix()
    cls = la.TDMA if 'TDMA' in name else la.PDMA
    f = torch.from_numpy(g[f'f_{name}_{axis}']).cuda()
    u = cls(M)(f, axis=axis)
    assert rel_l2(u.cpu().numpy(), g[f'u_{name}_{axis}']) < TOL


def _poisson_space(fam, n=14):
    SD = sf.FunctionSpace(n, fam, bc=(0, 0))
    K1 = sf.FunctionSpace(12, 'F', dtype='D')
    K2 = sf.FunctionSpace(10, 'F', dtype='d')
    return sf.TensorProductSpace(sf.comm, (SD, K1, K2))


def test_solver_generic_1nd_poisson_and_helmholtz_legendre():
    g = load_golden('solvers')
    T = _poisson_space('L')
    u, v = sf.TrialFunction(T), sf.TestFunction(T)
    f_hat = sf.Function(T, buffer=g['poisson_leg_f'])
    H = la.SolverGeneric1ND(sf.inner(v, sf.div(sf.grad(u))))
    u_hat = H(f_hat, sf.Function(T), constraints=((0, 0),))
    assert rel_l2(u_hat.get(), g['poisson_leg_u']) < TOL
    H2 = la.SolverGeneric1ND(sf.inner(v, sf.div(sf.grad(u))-2.5*u))
    u2 = H2(f_hat, sf.Function(T))
    assert rel_l2(u2.get(), g['helm_leg_u']) < TOL
    # solve o matvec round trip (tests/test_matrices.py:559-591)
    c = H2.matvec(u2, sf.Function(T))
    d = T.bases[0].dim()
    assert rel_l2(c.get()[:d], g['poisson_leg_f'][:d]) < TOL


def test_chebyshev_helmholtz_golden():
    g = load_golden('solvers')
    T = _poisson_space('C')
    u, v = sf.TrialFunction(T), sf.TestFunction(T)
    f_hat = sf.Function(T, buffer=g['poisson_cheb_f'])
    H = la.Helmholtz(sf.inner(v, sf.div(sf.grad(u))-2.5*u))
    u_hat = H(f_hat, sf.Function(T))
    assert rel_l2(u_hat.get(), g['helm_cheb_u']) < TOL
    c = H.matvec(u_hat, sf.Function(T))
    assert rel_l2(c.get(), g['helm_cheb_matvec']) < 1e-12
    d = T.bases[0].dim()
    assert rel_l2(c.get()[:d], g['poisson_cheb_f'][:d]) < 1e-12


def test_biharmonic_legendre_golden():
    g = load_golden('solvers')
    SB = sf.FunctionSpace(16, 'L', bc=(0, 0, 0, 0))
    T = sf.TensorProductSpace(sf.comm, (SB, sf.FunctionSpace(12, 'F', dtype='D'), sf.FunctionSpace(10, 'F', dtype='d')))
    u, v = sf.TrialFunction(T), sf.TestFunction(T)
    H = la.SolverGeneric1ND(sf.inner(v, sf.div(sf.grad(sf.div(sf.grad(u))))))
    u_hat = H(sf.Function(T, buffer=g['biharm_f']), sf.Function(T))
    assert rel_l2(u_hat.get(), g['biharm_u']) < 1e-12


def test_gemm_axis_shapes():
    rng = np.random.default_rng(2)
    for (M, K, outer, inner, cplx) in [(37, 29, 3, 11, True), (64, 64, 1, 130, False), (70, 50, 9, 1, False),
                                        (33, 65, 5, 1, True), (128, 96, 2, 257, True)]:
        A = rng.standard_normal((M, K))
        u = rng.standard_normal((outer, K, inner))
        if cplx:
            u = u+1j*rng.standard_normal(u.shape)
        out = torch.empty((outer, M, inner), dtype=torch.complex128 if cplx else torch.float64, device='cuda')
        sf.ops.gemm_axis(torch.from_numpy(A).cuda(), torch.from_numpy(u).cuda(), out, outer, inner)
        ref = np.einsum('ij,ojc->oic', A, u)
        assert rel_l2(out.cpu().numpy(), ref) < 1e-14


@pytest.mark.parametrize('pf', [1.0, 1.5])
def test_fast_dct_sizes_complex_and_padded(pf):
    """Sizes served by the register-resident DCT kernels (64..768), plain and 3/2-padded."""
    rng = np.random.default_rng(8)
    for N, name in ((64, 'dirichlet'), (128, 'biharmonic'), (256, 'neumann'), (512, 'dirichlet'), (96, 'ortho')):
        S = sf.FunctionSpace(N, 'C', bc=BCS[name], padding_factor=pf)
        spec = oracle_spec(S)
        Np = S.physical_len()
        u = rng.standard_normal((Np, 3, 7))+1j*rng.standard_normal((Np, 3, 7))
        out = torch.empty((N, 3, 7), dtype=torch.complex128, device='cuda')
        S._scalar_product_axis(torch.from_numpy(u).cuda(), out, 1, 21)
        d = S.dim()
        assert rel_l2(out.cpu().numpy()[:d], opb.scalar_product_axis(u, 0, spec)[:d]) < TOL, (N, name)
        assert np.all(out.cpu().numpy()[d:] == 0)
        vh = rng.standard_normal((N, 3, 7))+1j*rng.standard_normal((N, 3, 7))
        vh[d:] = 0
        outp = torch.empty((Np, 3, 7), dtype=torch.complex128, device='cuda')
        S._backward_axis(torch.from_numpy(vh).cuda(), outp, 1, 21)
        assert rel_l2(outp.cpu().numpy(), opb.backward_axis(vh, 0, spec)) < TOL, (N, name)


def test_all_chebyshev_real_2d():
    """Real data through the Chebyshev kernels (strided fast path on axis 0, generic on axis 1)."""
    rng = np.random.default_rng(9)
    S0 = sf.FunctionSpace(64, 'C', bc=(0, 0))
    S1 = sf.FunctionSpace(24, 'C', bc=(0, 0, 0, 0))
    T = sf.TensorProductSpace(sf.comm, (S0, S1))
    u = rng.standard_normal((64, 24))
    uh = T.forward(u).get()
    ref = opb.forward_axis(opb.forward_axis(u, 1, oracle_spec(T.bases[1])), 0, oracle_spec(T.bases[0]))
    assert uh.dtype == np.float64
    assert rel_l2(uh, ref) < 1e-12
    ub = T.backward(sf.Function(T, buffer=ref)).get()
    refb = opb.backward_axis(opb.backward_axis(ref, 0, oracle_spec(T.bases[0])), 1, oracle_spec(T.bases[1]))
    assert rel_l2(ub, refb) < TOL


def test_chebyshev_biharmonic_solver_golden():
    """chebyshev.la.Biharmonic (a18): device LU with the rank-2 tail vs the reference's solver."""
    g = load_golden('solvers')
    SB = sf.FunctionSpace(16, 'C', bc=(0, 0, 0, 0))
    T = sf.TensorProductSpace(sf.comm, (SB, sf.FunctionSpace(12, 'F', dtype='D'), sf.FunctionSpace(10, 'F', dtype='d')))
    u, v = sf.TrialFunction(T), sf.TestFunction(T)
    H = la.Biharmonic(sf.inner(v, sf.div(sf.grad(sf.div(sf.grad(u))))))
    f_hat = sf.Function(T, buffer=g['cheb_biharm_f'])
    u_hat = H(f_hat, sf.Function(T))
    # The device LU (O(n) with the rank-2 tail carried along) is not the reference's elimination order
    # (la.pyx:975-1127): besides the 1e-12 agreement with the reference's output, both are checked
    # against the solution of the same (fp64-entry) systems in 40-digit arithmetic.
    import mpmath as mp
    mp.mp.dps = 40
    d = SB.dim()
    S4 = SB.spectral_matrix(0, 4).diags()[:d, :d]
    A2 = SB.spectral_matrix(0, 2).diags()[:d, :d]
    Bm = SB.mass_matrix().diags()[:d, :d]
    k1 = np.fft.fftfreq(12, 1./12)
    k2 = np.fft.rfftfreq(10, 1./10)
    xd, xr, f = u_hat.get(), g['cheb_biharm_u'], g['cheb_biharm_f']
    num_d = num_r = den = mp.mpf(0)
    for j in (0, 1, 5, 6, 11):                       # a sample of lines incl. the largest wavenumbers
        for l in (0, 2, 5):
            kk = float(k1[j]**2+k2[l]**2)
            Hm = mp.matrix((S4-2*kk*A2+kk**2*Bm).tolist())
            for part in ('real', 'imag'):
                rhs = mp.matrix([mp.mpf(float(getattr(v, part))) for v in f[:d, j, l]])
                xe = mp.lu_solve(Hm, rhs)
                for i in range(d):
                    num_d += (mp.mpf(float(getattr(xd[i, j, l], part)))-xe[i])**2
                    num_r += (mp.mpf(float(getattr(xr[i, j, l], part)))-xe[i])**2
                    den += xe[i]**2
    e_dev, e_ref = float(mp.sqrt(num_d/den)), float(mp.sqrt(num_r/den))
    print(f'chebyshev biharmonic vs 40-digit solve: device {e_dev:.2e}, reference {e_ref:.2e}')
    assert e_dev <= max(3*e_ref, TOL), (e_dev, e_ref)
    assert rel_l2(xd, xr) < TOL
    c = H.matvec(u_hat, sf.Function(T))
    d = SB.dim()
    assert rel_l2(c.get()[:d], g['cheb_biharm_f'][:d]) < 1e-12
    cref = H.matvec(sf.Function(T, buffer=g['cheb_biharm_u']), sf.Function(T))
    assert rel_l2(cref.get(), g['cheb_biharm_matvec']) < 1e-12


def test_chebyshev_biharmonic_larger_residual():
    rng = np.random.default_rng(4)
    SB = sf.FunctionSpace(128, 'C', bc=(0, 0, 0, 0))
    T = sf.TensorProductSpace(sf.comm, (SB, sf.FunctionSpace(16, 'F', dtype='D'), sf.FunctionSpace(12, 'F', dtype='d')))
    u, v = sf.TrialFunction(T), sf.TestFunction(T)
    H = la.Biharmonic(sf.inner(v, sf.div(sf.grad(sf.div(sf.grad(u))))))
    f = rng.standard_normal(T.shape(True))+1j*rng.standard_normal(T.shape(True))
    f[SB.dim():] = 0
    x = H(sf.Function(T, buffer=f), sf.Function(T))
    c = H.matvec(x, sf.Function(T))
    assert rel_l2(c.get()[:SB.dim()], f[:SB.dim()]) < 1e-12


def test_chebyshev_gauss_lobatto_in_a_tensor_product_space():
    """quad='GL' along one axis of a 3-D space: forward o backward is the identity on the dofs and
    the Lobatto points include both end points."""
    SD = sf.FunctionSpace(20, 'C', bc=(0, 0), quad='GL')
    K1 = sf.FunctionSpace(12, 'F', dtype='D')
    K2 = sf.FunctionSpace(10, 'F', dtype='d')
    T = sf.TensorProductSpace(sf.comm, (SD, K1, K2))
    x = SD.points_and_weights()[0]
    assert abs(x[0]-1) < 1e-15 and abs(x[-1]+1) < 1e-15
    rng = np.random.default_rng(11)
    vh = np.zeros(T.shape(True), dtype=complex)
    vh[:SD.dim()] = rng.standard_normal((SD.dim(),)+tuple(T.shape(True)[1:]))
    vh[:, :, 0] = vh[:, :, 0].real
    u = T.backward(sf.Function(T, buffer=vh))
    back = T.forward(T.backward(T.forward(u)))
    assert rel_l2(back.get(), T.forward(u).get()) < TOL


@pytest.mark.parametrize('name', ['CTSD', 'CTSB', 'CTT', 'CTSN', 'ASDSD'])
@pytest.mark.parametrize('axis', [0, 1, 2])
def test_separable_tail_matvec(name, axis):
    """Upper-triangular Chebyshev derivative matrices applied as banded part + running sums
    (`la.separable_tail`, `shb_tail_matvec`) == the dense product."""
    from shenfun_b200 import ops
    N = 40
    TC = sf.FunctionSpace(N, 'C')
    SD = sf.FunctionSpace(N, 'C', bc=(0, 0))
    SB = sf.FunctionSpace(N, 'C', bc=(0, 0, 0, 0))
    SN = sf.FunctionSpace(N, 'C', bc=BCS['neumann'])
    test, trial, a, b = {'CTSD': (TC, SD, 0, 1), 'CTSB': (TC, SB, 0, 1), 'CTT': (TC, TC, 0, 1), 'CTSN': (TC, SN, 0, 1),
                         'ASDSD': (SD, SD, 0, 2)}[name]
    D = test.spectral_matrix(a, b, trial).diags()
    w, r, t, band = la.separable_tail(D)
    n_rows, n_cols = D.shape
    rng = np.random.default_rng(4)
    shape = [3, 4, 5]
    shape[axis] = N
    u = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
    sl = [slice(None)]*3
    sl[axis] = slice(0, n_cols)
    ref = np.zeros(shape, dtype=complex)
    so = [slice(None)]*3
    so[axis] = slice(0, n_rows)
    ref[tuple(so)] = np.moveaxis(np.tensordot(D, np.moveaxis(u[tuple(sl)], axis, 0), axes=(1, 0)), 0, axis)
    outer = int(np.prod(shape[:axis]))
    inner = int(np.prod(shape[axis+1:]))
    scale = rng.uniform(0.5, 2.0, outer*inner)
    sshape = list(shape)
    sshape[axis] = 1
    ref = ref*scale.reshape(sshape)
    ud = torch.from_numpy(u).cuda()
    out = torch.full(shape, 7.0, dtype=torch.complex128, device='cuda')
    sc = torch.from_numpy(scale).cuda()
    acc = 0
    if band:
        offs = sorted(band)
        tab = np.zeros((len(offs), n_rows))
        for k, o in enumerate(offs):
            v = band[o]
            if o >= 0:
                tab[k, :len(v)] = v
            else:
                tab[k, -o:-o+len(v)] = v
        ops.banded_matvec(ud, out, offs, [torch.from_numpy(tab).cuda()], [sc], outer, inner, n_rows, n_cols, N, N, 0)
        acc = 1
    ops.tail_matvec(ud, out, torch.from_numpy(r).cuda(), torch.from_numpy(t).cuda(), sc, w, outer, inner, n_rows,
                    n_cols, N, N, acc)
    assert rel_l2(out.cpu().numpy(), ref) < TOL
    # times-i flag on top of an existing array
    out2 = torch.from_numpy(ref.copy()).cuda()
    ops.tail_matvec(ud, out2, torch.from_numpy(r).cuda(), torch.from_numpy(t).cuda(), sc, w, outer, inner, n_rows,
                    n_cols, N, N, 3)
    tail_only = ref.copy()
    if band:
        tb = torch.zeros(shape, dtype=torch.complex128, device='cuda')
        ops.banded_matvec(ud, tb, offs, [torch.from_numpy(tab).cuda()], [sc], outer, inner, n_rows, n_cols, N, N, 0)
        tail_only = ref-tb.cpu().numpy()
    assert rel_l2(out2.cpu().numpy(), ref+1j*tail_only) < TOL


@pytest.mark.parametrize('fam', ['C', 'L'])
def test_poisson3d_demo_manufactured_solution(fam):
    """BASELINE config 2 (demo/poisson3D.py structure) at 32^3: the device solution reproduces the
    manufactured solution to spectral accuracy (the reference demo asserts 1e-6)."""
    from shenfun_b200.demos import poisson3d
    err, _ = poisson3d.main((32, 32, 32), fam)
    assert err < 1e-8, err


@pytest.mark.parametrize('fam', ['L', 'C'])
def test_biharmonic3d_demo_manufactured_solution(fam):
    """BASELINE config 3 (demo/biharmonic3D.py structure) at 32^3."""
    from shenfun_b200.demos import biharmonic3d
    err, _ = biharmonic3d.main((32, 32, 32), fam)
    assert err < 1e-8, err


@pytest.mark.parametrize('fam', ['C', 'L'])
@pytest.mark.parametrize('layout', ['first', 'middle'])
@pytest.mark.parametrize('kind', ['const', 'func', 'neumann'])
def test_3d_inhomogeneous_boundary_values_golden(fam, layout, kind):
    """Inhomogeneous boundary values in a tensor-product space: the boundary data
    is a function of the other coordinates and is transformed along them
    (`BoundaryValues.set_tensor_bcs`, tensorproductspace.py:1909-2032); forward, backward and
    `Function.set_boundary_dofs` against the reference's own run (tests/golden/make_golden.py::poly_3d_inhom)."""
    import sympy as sp
    x, y, z = sp.symbols('x,y,z', real=True)
    g = load_golden('poly3d_inhom')
    other = (y, z) if layout == 'first' else (x, z)
    bc = {
        'const': {'left': {'D': 1.5}, 'right': {'D': -0.5}},
        'func': {'left': {'D': sp.cos(2*other[0])*sp.sin(other[1])}, 'right': {'D': 0.25+sp.sin(other[0])}},
        'neumann': {'left': {'N': 0.75}, 'right': {'N': sp.cos(other[1])}},
    }[kind]
    SD = sf.FunctionSpace(14, fam, bc=bc)
    K1 = sf.FunctionSpace(12, 'F', dtype='D')
    K2 = sf.FunctionSpace(10, 'F', dtype='d')
    T = sf.TensorProductSpace(sf.comm, (SD, K1, K2)) if layout == 'first' else \
        sf.TensorProductSpace(sf.comm, (K1, SD, K2), axes=(1, 0, 2))
    tag = f'{fam}_{kind}_{layout}'
    uh = T.forward(g[f'u_{tag}'])
    assert rel_l2(uh.get(), g[f'fw_{tag}']) < TOL
    assert rel_l2(T.backward(sf.Function(T, buffer=g[f'fw_{tag}'])).get(), g[f'bw_{tag}']) < TOL
    f0 = sf.Function(T)
    f0.set_boundary_dofs()
    assert rel_l2(f0.get(), g[f'bdofs_{tag}']) < TOL
