"""``WDBX`` facade for the search path (reference: wdbx/core/wdbx.py:19-502).

Same constructor arguments and the same ``vector_search / vector_search_async /
vector_store_async / get_vector / delete_vector / update_metadata / count_vectors / clear /
get_stats / initialize / shutdown`` methods, result shapes and ``ValueError("Vector dimension
mismatch ...")`` behaviour (wdbx.py:323-326).  Plugins, REST, CLI and the TCP shard manager are
callers / control plane and deliberately not part of this package:
``enable_plugins`` and ``enable_distributed`` are accepted and ignored.

This package IS the ``enable_gpu=True`` path that the reference leaves dead (wdbx.py:120-126 ->
vector_store.py:124-130 never consume the flag); ``enable_gpu=False`` raises, there is no CPU
fallback.
"""
from __future__ import annotations

import logging
import uuid
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

from . import __version__ as _pkg_version  # noqa: F401  (set in __init__ before this import)
from .config import WDBXConfig
from .vector_store import BatchResult, VectorStore

logger = logging.getLogger(__name__)


class WDBX:
    def __init__(
        self,
        vector_dimension: int = 384,
        num_shards: int = 1,
        data_dir: str = "./wdbx_data",
        config: Optional[Dict[str, Any]] = None,
        enable_plugins: bool = False,
        enable_distributed: bool = False,
        enable_gpu: bool = True,
        log_level: str = "INFO",
    ):
        numeric_level = getattr(logging, str(log_level).upper(), None)
        if not isinstance(numeric_level, int):
            raise ValueError(f"Invalid log level: {log_level}")
        logger.setLevel(numeric_level)
        self.vector_dim = vector_dimension
        self.num_shards = num_shards
        self.data_dir = Path(data_dir)
        self.config = config if isinstance(config, WDBXConfig) else WDBXConfig(config or {})
        self.enable_plugins = False
        self.enable_distributed = enable_distributed
        self.enable_gpu = enable_gpu
        self.plugins: Dict[str, Any] = {}
        self.shard_manager = None
        if not self.data_dir.exists():
            self.data_dir.mkdir(parents=True)
        self._init_vector_store()

    @property
    def version(self) -> str:
        from . import __version__

        return __version__

    def _init_vector_store(self):
        """wdbx.py:116-126 -- here the flag finally selects an engine."""
        self._store = VectorStore(
            vector_dim=self.vector_dim,
            data_dir=self.data_dir,
            num_shards=self.num_shards,
            use_gpu=self.enable_gpu,
            config=self.config,
        )

    # The reference names both an attribute and a method ``vector_store`` (wdbx.py:120 / :241); the
    # attribute wins at run time, so ``db.vector_store`` is the VectorStore object there.  Keep that.
    @property
    def vector_store(self) -> VectorStore:
        return self._store

    async def initialize(self):
        await self._store.initialize()

    async def shutdown(self):
        await self._store.shutdown()

    def close(self):
        self._store.close()

    # ------------------------------------------------------------------ store
    def _check_dim(self, vector):
        if len(vector) != self.vector_dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.vector_dim}, got {len(vector)}")

    def store_vector(self, vector: List[float], metadata: Optional[Dict[str, Any]] = None,
                     id: Optional[str] = None) -> str:
        """Sync counterpart of ``vector_store_async`` (the reference's sync method is shadowed, wdbx.py:241)."""
        self._check_dim(vector)
        vector_id = id or str(uuid.uuid4())
        self._store.store(vector_id, vector, metadata)
        return vector_id

    async def vector_store_async(self, vector: List[float], metadata: Optional[Dict[str, Any]] = None,
                                 id: Optional[str] = None) -> str:
        """wdbx.py:272-301."""
        self._check_dim(vector)
        vector_id = id or str(uuid.uuid4())
        await self._store.store_async(vector_id, vector, metadata)
        return vector_id

    # ------------------------------------------------------------------ search
    def vector_search(self, query_vector: List[float], limit: int = 10, threshold: float = 0.0,
                      filter_metadata: Optional[Dict[str, Any]] = None) -> List[Tuple[str, float, Dict[str, Any]]]:
        """wdbx.py:303-336."""
        self._check_dim(query_vector)
        return self._store.search(query_vector, limit=limit, threshold=threshold, filter_metadata=filter_metadata)

    async def vector_search_async(self, query_vector: List[float], limit: int = 10, threshold: float = 0.0,
                                  filter_metadata: Optional[Dict[str, Any]] = None):
        """wdbx.py:338-371."""
        self._check_dim(query_vector)
        return await self._store.search_async(query_vector, limit=limit, threshold=threshold,
                                              filter_metadata=filter_metadata)

    def vector_search_batch(self, queries, limit: int = 10, metric: Optional[str] = None) -> BatchResult:
        """Additive batch API: [B, dim] queries in one device call."""
        return self._store.search_batch(queries, limit=limit, metric=metric)

    # ------------------------------------------------------------------ CRUD passthroughs (wdbx.py:373-470)
    def delete_vector(self, vector_id: str) -> bool:
        return self._store.delete(vector_id)

    async def delete_vector_async(self, vector_id: str) -> bool:
        return await self._store.delete_async(vector_id)

    def update_metadata(self, vector_id: str, metadata: Dict[str, Any]) -> bool:
        return self._store.update_metadata(vector_id, metadata)

    async def update_metadata_async(self, vector_id: str, metadata: Dict[str, Any]) -> bool:
        return await self._store.update_metadata_async(vector_id, metadata)

    def get_vector(self, vector_id: str):
        return self._store.get(vector_id)

    async def get_vector_async(self, vector_id: str):
        return await self._store.get_async(vector_id)

    def count_vectors(self) -> int:
        return self._store.count()

    def clear(self) -> int:
        return self._store.clear()

    async def clear_async(self) -> int:
        return await self._store.clear_async()

    def get_stats(self) -> Dict[str, Any]:
        """wdbx.py:480-502."""
        stats = {
            "version": self.version,
            "vector_dimension": self.vector_dim,
            "num_shards": self.num_shards,
            "total_vectors": self.count_vectors(),
            "plugins_enabled": self.enable_plugins,
            "plugins_loaded": 0,
            "distributed_enabled": self.enable_distributed,
            "gpu_enabled": self.enable_gpu,
        }
        stats.update(self._store.get_stats())
        return stats
