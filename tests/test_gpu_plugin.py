"""The by-name kernel plugin (shenfun_b200/optimization) against outputs of the reference's
own numba kernels called with the same positional arguments (tests/golden/plugin.npz, written
by tests/golden/make_golden.py::plugin_kernels).  Tolerance 1e-12 relative L2 (fp64)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-12
KINDS = ('TDMA', 'TDMA_O', 'PDMA', 'FDMA', 'TwoDMA', 'ThreeDMA', 'HeptaDMA', 'DiagMA')


@pytest.fixture(scope='module')
def g():
    return load_golden('plugin')


@pytest.fixture(scope='module')
def b200():
    from shenfun_b200.optimization import b200
    return b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class _Mat(dict):
    scale = 1.0


@pytest.mark.parametrize('kind', KINDS)
def test_banded_lu_and_solve(g, b200, kind):
    data = dev(g[f'{kind}_data'])
    if hasattr(b200, f'{kind}_LU'):
        getattr(b200, f'{kind}_LU')(data)
        # only the entries inside the matrix are defined (the DIA corners are never read)
        ref = g[f'{kind}_lu']
        n = ref.shape[1]
        from shenfun_b200.optimization.b200.la import _offsets
        for d, off in enumerate(_offsets(kind)):
            sl = slice(off, n) if off >= 0 else slice(0, n+off)
            assert rel_l2(data[d, sl].cpu().numpy(), ref[d, sl]) < TOL, (kind, off)
    lu = dev(g[f'{kind}_lu'])
    if kind == 'ThreeDMA':
        # All three reference implementations read the +4 diagonal from row 1 of `data`
        # (`u2 = data[1, 4:]`: la.py:697, numba/threedma.py:28, la.pyx:574), so their result is not
        # the solution of the stored matrix.  The device kernel solves the matrix that is stored;
        # check it against a dense solve instead of the reference output.
        n = lu.shape[1]
        A = sum(np.diag(g['ThreeDMA_data'][d, off:], off) for d, off in enumerate((0, 2, 4)))
        dense = lambda b, axis: np.moveaxis(np.linalg.solve(A, np.moveaxis(b, axis, 0)[:n].reshape(n, -1)).reshape(
            np.moveaxis(b, axis, 0)[:n].shape), 0, axis)
    for axis in (0, 1, 2):
        u = dev(g[f'{kind}_b{axis}'])
        r = getattr(b200, f'{kind}_Solve')(u, lu, axis)
        assert r is u
        ref = g[f'{kind}_u{axis}'] if kind != 'ThreeDMA' else dense(g[f'{kind}_b{axis}'], axis)
        assert rel_l2(u.cpu().numpy(), ref) < TOL, (kind, axis)
    u = dev(g[f'{kind}_b1d'])
    getattr(b200, f'{kind}_inner_solve')(u, lu)
    ref = g[f'{kind}_u1d'] if kind != 'ThreeDMA' else np.concatenate([dense(g[f'{kind}_b1d'], 0), g[f'{kind}_b1d'][n:]])
    assert rel_l2(u.cpu().numpy(), ref) < TOL


def test_banded_numpy_arguments_are_staged_in_place(g, b200):
    """The reference's own (host) arrays: staged to the device and written back in place."""
    data = g['PDMA_data'].copy()
    b200.PDMA_LU(data)
    u = g['PDMA_b1'].copy()
    b200.PDMA_Solve(u, data, 1)
    assert rel_l2(u, g['PDMA_u1']) < TOL


@pytest.mark.parametrize('kind', ('TDMA', 'PDMA', 'FDMA'))
@pytest.mark.parametrize('naxes', (0, 1, 2))
def test_solver_generic_1nd_solve_data(g, b200, kind, naxes):
    data = dev(g[f'g1nd_{kind}_{naxes}_data'])
    u = dev(g[f'g1nd_{kind}_{naxes}_b'])
    b200.SolverGeneric1ND_solve_data(u, data, getattr(b200, f'{kind}_inner_solve'), naxes, True)
    assert rel_l2(u.cpu().numpy(), g[f'g1nd_{kind}_{naxes}_u']) < TOL


@pytest.mark.parametrize('tag', ('D', 'N'))
@pytest.mark.parametrize('axis', (0, 1, 2))
def test_helmholtz(g, b200, tag, axis):
    A = _Mat({k: g[f'helm{tag}_A{k}'] for k in (0, 2, 4)})
    B = _Mat({k: g[f'helm{tag}_B{k}'] for k in (-2, 0, 2)})
    neumann = tag == 'N'
    shape = g[f'helm{tag}_{axis}_d0'].shape
    d0, d1, d2, L = (torch.zeros(shape, dtype=torch.float64, device='cuda') for _ in range(4))
    alfa, beta = g[f'helm{tag}_{axis}_alfa'], g[f'helm{tag}_{axis}_beta']
    b200.LU_Helmholtz(A, B, alfa, beta, neumann, d0, d1, d2, L, axis)
    n = shape[axis]-2
    idx = [slice(None)]*3

    def rows(t, upto):
        idx[axis] = slice(0, upto)
        return (t.cpu().numpy() if isinstance(t, torch.Tensor) else t)[tuple(idx)]
    for nm, t, upto in (('d0', d0, n), ('d1', d1, n-2), ('d2', d2, n-4), ('L', L, n-2)):
        assert rel_l2(rows(t, upto), rows(g[f'helm{tag}_{axis}_{nm}'], upto)) < TOL, nm
    b = dev(g[f'helm{tag}_{axis}_b'])
    u = torch.zeros_like(b)
    b200.Solve_Helmholtz(b, u, neumann, d0, d1, d2, L, axis)
    assert rel_l2(rows(u, n), rows(g[f'helm{tag}_{axis}_u'], n)) < TOL
    if not neumann:
        c = torch.zeros_like(b)
        b200.Helmholtz_matvec(dev(g[f'helm{tag}_{axis}_u']), c, alfa, beta, A, B, axis)
        assert rel_l2(rows(c, n), rows(g[f'helm{tag}_{axis}_matvec'], n)) < TOL


@pytest.mark.parametrize('axis', (0, 1, 2))
def test_poisson_solve_add(g, b200, axis):
    A = _Mat({0: g['add_A0'], 2: g['add_A2']})
    A.scale = float(g['add_scale'])
    b = dev(g[f'add_{axis}_b'])
    u = torch.zeros_like(b)
    b200.Poisson_Solve_ADD(A, b, u, axis)
    n = g['add_A0'].shape[0]
    idx = [slice(None)]*3
    idx[axis] = slice(0, n)
    assert rel_l2(u.cpu().numpy()[tuple(idx)], g[f'add_{axis}_u'][tuple(idx)]) < TOL


@pytest.mark.parametrize('axis', (0, 1, 2))
def test_biharmonic(g, b200, axis):
    names = ('sii', 'siu', 'siuu', 'ail', 'aii', 'aiu', 'bill', 'bil', 'bii', 'biu', 'biuu')
    dg = [g[f'bih_{k}'] for k in names]
    a0 = float(g['bih_a0'])
    alfa, beta = g[f'bih_{axis}_alfa'], g[f'bih_{axis}_beta']
    fshape = g[f'bih_{axis}_u0'].shape
    u0, u1, u2, l0, l1, ak, bk = (torch.zeros(fshape, dtype=torch.float64, device='cuda') for _ in range(7))
    b200.LU_Biharmonic(a0, alfa, beta, *dg, u0, u1, u2, l0, l1, axis)
    b200.Biharmonic_factor_pr(ak, bk, l0, l1, axis)
    M = fshape[axis+1]
    idx = [slice(None)]*4

    def rows(t, upto):
        idx[axis+1] = slice(0, upto)
        return (t.cpu().numpy() if isinstance(t, torch.Tensor) else t)[tuple(idx)]
    for nm, t, upto in (('u0', u0, M), ('u1', u1, M-1), ('u2', u2, M-2), ('l0', l0, M-1), ('l1', l1, M-2),
                        ('ak', ak, M-3), ('bk', bk, M-3)):
        assert rel_l2(rows(t, upto), rows(g[f'bih_{axis}_{nm}'], upto)) < 1e-12, nm
    b = dev(g[f'bih_{axis}_b'])
    u = torch.zeros_like(b)
    b200.Biharmonic_Solve(b, u, u0, u1, u2, l0, l1, ak, bk, a0, axis)
    n = 2*M
    i3 = [slice(None)]*3
    i3[axis] = slice(0, n)
    assert rel_l2(u.cpu().numpy()[tuple(i3)], g[f'bih_{axis}_u'][tuple(i3)]) < 1e-12
    # second solve hits the packed-factor cache; a change of the factors invalidates it
    u2nd = torch.zeros_like(b)
    b200.Biharmonic_Solve(b, u2nd, u0, u1, u2, l0, l1, ak, bk, a0, axis)
    assert torch.equal(u, u2nd)
    # from the reference's own factor arrays (host arrays)
    uh = np.zeros_like(g[f'bih_{axis}_b'])
    b200.Biharmonic_Solve(g[f'bih_{axis}_b'].copy(), uh, *[g[f'bih_{axis}_{k}'] for k in
                                                          ('u0', 'u1', 'u2', 'l0', 'l1', 'ak', 'bk')], a0, axis)
    assert rel_l2(uh[tuple(i3)], g[f'bih_{axis}_u'][tuple(i3)]) < 1e-12
    c = torch.zeros_like(b)
    b200.Biharmonic_matvec(dev(g[f'bih_{axis}_u']), c, a0, alfa, beta, *dg, axis)
    assert rel_l2(c.cpu().numpy()[tuple(i3)], g[f'bih_{axis}_matvec'][tuple(i3)]) < 1e-12


@pytest.mark.parametrize('tag', ('leg', 'jac', 'chebu'))
@pytest.mark.parametrize('axis', (0, 1, 2))
@pytest.mark.parametrize('dt', ('real', 'cplx'))
def test_recursive_transforms(g, b200, tag, axis, dt):
    x, w, a = g[f'tr_{tag}_x'], g[f'tr_{tag}_w'], g[f'tr_{tag}_a']
    uin = dev(g[f'tr_{tag}_{axis}_{dt}_in'])
    sp = torch.zeros_like(uin)
    b200.transforms.scalar_product(uin, sp, x, w, axis, a)
    assert rel_l2(sp.cpu().numpy(), g[f'tr_{tag}_{axis}_{dt}_sp']) < TOL
    ev = torch.zeros_like(uin)
    b200.transforms.evaluate_expansion_all(dev(g[f'tr_{tag}_{axis}_{dt}_sp']), ev, x, axis, a)
    assert rel_l2(ev.cpu().numpy(), g[f'tr_{tag}_{axis}_{dt}_ev']) < TOL


def test_matvec_banded(b200):
    rng = np.random.default_rng(5)
    n = 17
    for axis in (0, 1, 2):
        shape = [4, 5, 6]
        shape[axis] = n
        v = rng.standard_normal(shape)+1j*rng.standard_normal(shape)
        diags = {o: rng.standard_normal(n-abs(o)) for o in (-4, -2, 0, 2, 4)}
        for offs, fn in (((-2, 0, 2), b200.Tridiagonal_matvec), ((-4, -2, 0, 2, 4), b200.Pentadiagonal_matvec)):
            Mx = sum(np.diag(diags[o], o) for o in offs)
            ref = np.moveaxis(np.tensordot(Mx, np.moveaxis(v, axis, 0), axes=(1, 0)), 0, axis)
            c = torch.zeros(shape, dtype=torch.complex128, device='cuda')
            fn(dev(v), c, axis, *[diags[o] for o in offs])
            assert rel_l2(c.cpu().numpy(), ref) < TOL


def test_pointwise(b200):
    rng = np.random.default_rng(6)
    a = rng.standard_normal((3, 6, 7, 8))
    b = rng.standard_normal((3, 6, 7, 8))
    c = torch.zeros((3, 6, 7, 8), dtype=torch.float64, device='cuda')
    b200.cross3D(c, dev(a), dev(b))
    assert rel_l2(c.cpu().numpy(), np.cross(a, b, axis=0)) < 1e-14
    c2 = torch.zeros((6, 7), dtype=torch.float64, device='cuda')
    b200.cross2D(c2, dev(a[:2, :, :, 0]), dev(b[:2, :, :, 0]))
    assert rel_l2(c2.cpu().numpy(), a[0, :, :, 0]*b[1, :, :, 0]-a[1, :, :, 0]*b[0, :, :, 0]) < 1e-14
    o = torch.zeros((9, 6, 7, 8), dtype=torch.float64, device='cuda')
    b200.outer3D(dev(a), dev(b), o, False)
    assert rel_l2(o.cpu().numpy(), (a[:, None]*b[None, :]).reshape(9, 6, 7, 8)) < 1e-14
    b200.outer3D(dev(a), dev(a), o, True)
    assert rel_l2(o.cpu().numpy(), (a[:, None]*a[None, :]).reshape(9, 6, 7, 8)) < 1e-14
    uh = rng.standard_normal((3, 6, 7, 5))+1j*rng.standard_normal((3, 6, 7, 5))
    mask = (rng.uniform(size=(6, 7, 5)) > 0.3).astype(np.int64)
    ud = dev(uh)
    b200.apply_mask(ud, mask)
    assert np.array_equal(ud.cpu().numpy(), uh*mask[None])
    mk = np.ones((6, 1, 1))
    mk[3] = 0
    ud = dev(uh[0])
    b200.apply_mask(ud, mk)
    assert np.array_equal(ud.cpu().numpy(), uh[0]*mk)
