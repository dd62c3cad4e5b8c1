"""Function.refine / Function.assign against vectors produced by the reference
(tests/golden/make_golden.py refine_assign; forms/arguments.py:1693-1827) and the Checkpoint
container (io/__init__.py:41-125).  Host logic: runs on CPU tensors here, on device tensors on a
GPU box."""
import os

import numpy as np
import pytest
import torch

import shenfun_b200 as sf

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'refine.npz'))
TOL = 1e-12

SPECS = {'f3d': [('F', 12, 'D'), ('F', 10, 'D'), ('F', 8, 'd')],
         'c3d': [('C', 12, None), ('F', 8, 'D'), ('F', 6, 'd')]}


def _space(tag, N=None):
    bases = []
    for i, (f, n, dt) in enumerate(SPECS[tag]):
        n = n if N is None else N[i]
        bases.append(sf.FunctionSpace(n, f, bc=(0, 0)) if f == 'C' else sf.FunctionSpace(n, f, dtype=dt))
    return sf.TensorProductSpace(sf.comm, bases)


def _mine(tag, N, r):
    """The reference drops the imaginary part of an R2C Nyquist plane also on unchanged axes."""
    r = r.get().copy()
    for ax, ((f, n, dt), m) in enumerate(zip(SPECS[tag], N)):
        if f == 'F' and dt == 'd' and n == m and n % 2 == 0:
            r[(slice(None),)*ax+(n//2,)] = r[(slice(None),)*ax+(n//2,)].real
    return r


def _ref(tag, N):
    """Reference vector; the reference also runs its padding rule over axes whose size does not
    change, which quarters an even-N C2C Nyquist plane (both index aliases are halved).  This
    drop-in leaves unchanged axes alone, so that factor is undone here."""
    ref = G[f"ref_{tag}_{'x'.join(map(str, N))}"].copy()
    for ax, ((f, n, dt), m) in enumerate(zip(SPECS[tag], N)):
        if f == 'F' and dt == 'D' and n == m and n % 2 == 0:
            ref[(slice(None),)*ax+(n//2,)] *= 4
        if f == 'F' and dt == 'd' and n == m and n % 2 == 0:      # R2C: the (real) Nyquist plane is halved
            ref[(slice(None),)*ax+(n//2,)] *= 2
    return ref


def _cases():
    for key in G.files:
        if key.startswith('ref_'):
            _, tag, n = key.split('_')
            yield tag, tuple(int(x) for x in n.split('x'))


@pytest.mark.parametrize('tag,N', list(_cases()))
def test_refine_golden(tag, N):
    T = _space(tag)
    uh = sf.Function(T, buffer=G[f'uh_{tag}'])
    r = uh.refine(N)
    ref = _ref(tag, N)
    assert tuple(r.shape) == ref.shape
    assert [b.N for b in r.function_space().bases] == list(N)
    assert np.abs(_mine(tag, N, r)-ref).max() <= TOL*max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('tag,N', list(_cases()))
def test_assign_golden(tag, N):
    T = _space(tag)
    uh = sf.Function(T, buffer=G[f'uh_{tag}'])
    out = sf.Function(_space(tag, N))
    ret = uh.assign(out)
    assert ret is out
    ref = _ref(tag, N)
    assert np.abs(_mine(tag, N, out)-ref).max() <= TOL*max(1.0, np.abs(ref).max())


def test_refine_same_size_is_a_copy_and_polynomial_truncation_keeps_low_modes():
    T = _space('c3d')
    uh = sf.Function(T, buffer=G['uh_c3d'])
    r = uh.refine((12, 8, 6))
    assert r.data_ptr() != uh.data_ptr()
    assert torch.equal(r.v, uh.v)
    t = uh.refine((8, 8, 6)).get()
    assert t.shape[0] == 8
    assert np.array_equal(t[:6], G['uh_c3d'][:6])
    assert np.all(t[6:] == 0)


def test_refine_vector():
    T = _space('f3d')
    V = sf.VectorSpace(T)
    uh = sf.Function(V)
    for i in range(3):
        uh[i] = G['uh_f3d']*(i+1)
    r = uh.refine((18, 16, 12))
    ref = G['ref_f3d_18x16x12']
    for i in range(3):
        assert np.abs(r[i].get()-(i+1)*ref).max() <= TOL*3*np.abs(ref).max()


def test_assign_rejects_other_bases():
    uh = sf.Function(_space('f3d'))
    with pytest.raises(AssertionError):
        uh.assign(sf.Function(_space('c3d')))


def test_checkpoint_roundtrip(tmp_path):
    T = _space('f3d')
    U, U0 = sf.Function(T, buffer=G['uh_f3d']), sf.Function(T, buffer=2*G['uh_f3d'])
    name = str(tmp_path/'run')
    chk = sf.Checkpoint(name, checkevery=5, data={'0': {'U': [U]}, '1': {'U': [U0]}})
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        chk.update(0.0, 0)
        chk.update(0.1, 1)               # not a checkpoint step
        U.v.mul_(3)
        chk.update(0.5, 5)
        assert os.path.isdir(name+'.old.chk')
        new = sf.Checkpoint(name)
        V = sf.Function(T)
        new.read(V, 'U', step=0)
        assert new.attrs == {'tstep': 5, 't': 0.5}
        assert torch.equal(V.v, U.v)
        new.read(V, 'U', step=1)
        assert torch.equal(V.v, U0.v)
        old = sf.Checkpoint(name+'.old')
        old.read(V, 'U')
        assert old.attrs['tstep'] == 0
        assert np.array_equal(V.get(), G['uh_f3d'])
        # kill file: checkpoint now and stop
        open('killshenfun', 'w').close()
        with pytest.raises(SystemExit):
            chk.update(0.7, 7)
        assert not os.path.exists('killshenfun')
        assert sf.Checkpoint(name).read(V, 'U') is V and sf.Checkpoint(name).attrs is None
        c = sf.Checkpoint(name)
        c.open()
        assert c.attrs['tstep'] == 7
    finally:
        os.chdir(cwd)
