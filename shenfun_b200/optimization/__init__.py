"""Kernel plugin resolved BY NAME -- the drop-in boundary of shenfun/optimization/__init__.py:47-84.

The reference's `optimizer(func)` looks `func.__name__`, then
`func.__qualname__.replace('.', '_')`, up in the module selected by
`config['optimization']['mode']` ('cython' or 'numba') and silently falls back to the Python
body when the name is missing (:72-77).  `b200` is the third such module: the same exported
names and positional signatures (optimization/cython/__init__.py:1-18), running on the device
through libshenb200.  Here a missing kernel is an error, never a fallback.
"""
from functools import wraps

from . import b200

__all__ = ('optimizer', 'runtimeoptimizer', 'b200')


def _resolve(func):
    fun = getattr(b200, func.__name__, None)
    if fun is None:
        fun = getattr(b200, func.__qualname__.replace('.', '_'), None)
    if fun is None:
        raise NotImplementedError(
            f'{func.__qualname__}: no b200 kernel of that name (shenfun_b200 never falls back to Python)')
    return fun


def optimizer(func, wrap=True):
    """Same contract as shenfun.optimization.optimizer, mode 'b200'."""
    fun = _resolve(func)
    if not wrap:
        return fun

    @wraps(func)
    def wrapped_function(*args, **kwargs):
        return fun(*args, **kwargs)
    return wrapped_function


class runtimeoptimizer:
    """Resolve at call time (shenfun/optimization/__init__.py:32-45)."""

    def __init__(self, func):
        self.func = func
        self.__doc__ = func.__doc__

    def __call__(self, *args, **kwargs):
        return _resolve(self.func)(*args, **kwargs)
