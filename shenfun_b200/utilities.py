"""Device versions of the helpers in shenfun/utilities/__init__.py that sit on
the Navier-Stokes path: `cross` (:177-210), `apply_mask` (:441-444), `dx` (:26-61)."""
import torch

from . import ops
from .arrays import Array, Function


def cross(c, a, b):
    """c = a x b for (3, ...) arrays; same argument order as the reference."""
    ops.cross3d(ops.plain(a), ops.plain(b), ops.plain(c))
    return c


def apply_mask(u_hat, mask):
    if mask is None:
        return u_hat
    t = ops.plain(u_hat)
    if not isinstance(mask, torch.Tensor):
        import numpy as np
        mask = ops.as_device_tensor(np.ascontiguousarray(
            np.broadcast_to(mask, t.shape[-mask.ndim:]), dtype=float))
    if mask.numel() != t[..., :].numel() and t.numel() % mask.numel() != 0:
        raise ValueError('mask shape does not match')
    if mask.shape != t.shape[-mask.dim():]:
        mask = mask.expand(t.shape[-mask.dim():]).contiguous()
    ops.apply_mask(t, mask.contiguous())
    return u_hat


def dx(u):
    """Integral of an Array over the reference domain: sum(u*w) with the un-weighted quadrature
    weights of every axis (utilities/__init__.py:26-61; "assumes a standard reference domain"), i.e.
    the true-domain integral of `forms._integrate` times the domain factors."""
    from .forms import _integrate
    space = u.function_space()
    df = 1.0
    for b in (space.flatten()[0].bases if space.is_composite_space else space.bases):
        df *= float(b.domain_factor())
    return _integrate(u)*df
