"""wdbx_b200 -- B200-native exact search behind WDBX's vector_search path.

Host-side mirror of the reference's operator interface for this path (same names, argument
meaning and error behaviour as wdbx/core/{wdbx,vector_store,indexing}.py) on top of
libwdbx_b200.so (hand-written sm_100a CUDA behind a C ABI, include/wdbx_b200.h).
There is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import B200Error, load_library  # noqa: E402,F401
from .config import WDBXConfig  # noqa: E402,F401
from .dist import DistContext  # noqa: E402,F401
from .engine import Engine, device_count  # noqa: E402,F401
from .indexing import B200FlatIndex, VectorIndex  # noqa: E402,F401
from .shard_map import ShardMap, shard_for_id, stable_hash  # noqa: E402,F401
from .vector_store import BatchResult, VectorStore  # noqa: E402,F401
from .wdbx import WDBX  # noqa: E402,F401

__all__ = ["B200Error", "load_library", "Engine", "device_count", "WDBX", "WDBXConfig", "VectorStore",
           "VectorIndex", "B200FlatIndex", "BatchResult", "DistContext", "ShardMap", "shard_for_id",
           "stable_hash"]
