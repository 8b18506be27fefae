"""wdbx_b200 -- B200-native exact search behind WDBX's vector_search path.

Host-side mirror of the reference's operator interface for this path (same names, argument
meaning and error behaviour as wdbx/core/{wdbx,vector_store,indexing}.py) on top of
libwdbx_b200.so (hand-written sm_100a CUDA behind a C ABI, include/wdbx_b200.h).
There is no CPU fallback.
"""
from ._lib import B200Error, load_library  # noqa: F401
from .engine import Engine, device_count  # noqa: F401

__version__ = "0.1.0"
__all__ = ["B200Error", "load_library", "Engine", "device_count"]
