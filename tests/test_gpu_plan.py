"""The C-level N-D plan API (`shb_plan_*`): one call per transform across the C ABI, against the
numpy oracle and the reference golden vectors.  Tolerance 1e-12 relative L2."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import tensorproduct as otp

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _descs(shape, kinds, padded):
    return [(k, n, p) for k, n, p in zip(kinds, shape, padded)]


@pytest.mark.parametrize('shape,kinds,pf,axes', [
    ((32, 32, 32), ('C2C', 'C2C', 'R2C'), 1.0, None),
    ((16, 16, 16), ('C2C', 'C2C', 'R2C'), 1.5, None),
    ((14, 15, 16), ('C2C', 'C2C', 'R2C'), 1.5, None),
    ((12, 10, 8), ('C2C', 'R2C', 'C2C'), 1.0, (0, 2, 1)),
    ((10, 12), ('C2C', 'C2C'), 2.0, (1, 0)),
    ((64,), ('R2C',), 1.5, None),
    ((6, 5, 4, 8), ('C2C', 'C2C', 'C2C', 'R2C'), 1.0, None),
])
def test_plan_against_oracle(shape, kinds, pf, axes):
    from shenfun_b200.plan import FourierPlan
    padded = [int(np.floor(pf*n)) for n in shape]
    plan = FourierPlan(shape, kinds, padded=padded, axes=axes)
    d = _descs(shape, kinds, padded)
    rng = np.random.default_rng(sum(shape))
    real = 'R2C' in kinds
    u = rng.standard_normal(padded) if real else rng.standard_normal(padded)+1j*rng.standard_normal(padded)
    uh = plan.forward(torch.from_numpy(u).cuda())
    assert rel_l2(uh.cpu().numpy(), otp.forward(u, d, axes=axes)) < TOL
    vh = rng.standard_normal(plan.spectral_shape)+1j*rng.standard_normal(plan.spectral_shape)
    v = plan.backward(torch.from_numpy(vh).cuda())
    assert rel_l2(v.cpu().numpy(), otp.backward(vh, d, axes=axes)) < TOL


def test_plan_against_reference_golden_and_batched_components():
    from shenfun_b200.plan import FourierPlan
    g = load_golden('f3d_14_15_16')
    N = tuple(int(n) for n in g['N'])
    plan = FourierPlan(N, ('C2C', 'C2C', 'R2C'))
    assert rel_l2(plan.forward(torch.from_numpy(g['u']).cuda()).cpu().numpy(), g['u_hat']) < TOL
    assert rel_l2(plan.backward(torch.from_numpy(g['v_hat']).cuda()).cpu().numpy(), g['v']) < TOL
    planp = FourierPlan(N, ('C2C', 'C2C', 'R2C'), padded=[int(1.5*n) for n in N])
    assert rel_l2(planp.backward(torch.from_numpy(g['v_hat']).cuda()).cpu().numpy(), g['v_pad15']) < TOL
    assert rel_l2(planp.forward(torch.from_numpy(g['w_pad15']).cuda()).cpu().numpy(), g['w_hat_pad15']) < TOL
    # three components in one call == three calls
    u3 = torch.randn((3,)+N, dtype=torch.float64, device='cuda')
    batched = plan.forward(u3, ncomp=3)
    for i in range(3):
        assert torch.equal(batched[i], plan.forward(u3[i].contiguous()))


def test_plan_argument_errors():
    from shenfun_b200 import _lib
    from shenfun_b200.plan import FourierPlan
    with pytest.raises(_lib.ShenB200Error):
        FourierPlan((8, 8), ('R2C', 'C2C'))                    # the real axis must be transformed first
    with pytest.raises(_lib.ShenB200Error):
        FourierPlan((8, 8), ('C2C', 'C2C'), axes=(0, 0))
    with pytest.raises(_lib.ShenB200Error):
        FourierPlan((8, 8), ('C2C', 'R2C'), padded=(12, 12), dealias_direct=True)
