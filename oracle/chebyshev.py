"""Oracle entry points used by oracle/tensorproduct.py for 'CHEB' axes (see polybases.py)."""
from .polybases import forward_axis, backward_axis, scalar_product_axis  # noqa: F401
