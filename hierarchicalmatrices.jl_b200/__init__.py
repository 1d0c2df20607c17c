"""B200-native hierarchical-matrix matvec engine behind the HierarchicalMatrices.jl API.

Only the `mul!` hot path lives here: the host-side mirror of the reference's
types (api.py), the ctypes binding of the C ABI (_lib.py), the CUDA sources
(csrc/) and the Julia `ccall` shim (julia/).  There is no CPU fallback.
"""
from . import _lib
from ._lib import HmError, build, lib
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all

__all__ = list(_api_all) + ["build", "lib", "HmError"]
