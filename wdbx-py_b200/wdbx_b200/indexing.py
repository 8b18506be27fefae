"""Operator boundary: the ``VectorIndex`` contract and its B200 implementation.

``VectorIndex`` restates the abstract interface of wdbx/core/indexing.py:18-217 (same method
names, argument meaning and return conventions).  ``B200FlatIndex`` is the per-shard facade the
store keeps in ``VectorStore.indices`` -- one per logical shard, as ``VectorStore._init_indices``
does (wdbx/core/vector_store.py:111-134) -- but all facades share ONE device engine, so a store
search is a single kernel launch instead of a Python loop over shards.
"""
from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Any, Dict, List, Tuple

import numpy as np

logger = logging.getLogger(__name__)


class VectorIndex(ABC):
    """Interface of a per-shard vector index (reference: indexing.py:18-217)."""

    @abstractmethod
    def __init__(self, vector_dim: int, index_path: Path, config: Any = None): ...

    @abstractmethod
    async def initialize(self): ...

    @abstractmethod
    async def shutdown(self): ...

    @abstractmethod
    def add(self, vector_id: str, vector: np.ndarray) -> bool: ...

    @abstractmethod
    async def add_async(self, vector_id: str, vector: np.ndarray) -> bool: ...

    @abstractmethod
    def batch_add(self, vectors: Dict[str, np.ndarray]) -> bool: ...

    @abstractmethod
    async def batch_add_async(self, vectors: Dict[str, np.ndarray]) -> bool: ...

    @abstractmethod
    def search(self, query_vector: np.ndarray, limit: int = 10) -> List[Tuple[str, float]]: ...

    @abstractmethod
    async def search_async(self, query_vector: np.ndarray, limit: int = 10) -> List[Tuple[str, float]]: ...

    @abstractmethod
    def remove(self, vector_id: str) -> bool: ...

    @abstractmethod
    async def remove_async(self, vector_id: str) -> bool: ...

    @abstractmethod
    def clear(self) -> bool: ...

    @abstractmethod
    async def clear_async(self) -> bool: ...

    @abstractmethod
    def optimize(self) -> bool: ...

    @abstractmethod
    async def optimize_async(self) -> bool: ...

    @abstractmethod
    def size(self) -> int: ...

    @abstractmethod
    def get_stats(self) -> Dict[str, Any]: ...


class B200FlatIndex(VectorIndex):
    """Exact flat index of ONE logical shard, backed by segment ``shard`` of the store's engine.

    Same conventions as ``FaissIndex`` (indexing.py:657-1183): ``search`` returns
    ``[(vector_id, similarity)]`` best-first with Python floats, ``limit`` is clipped to the
    live size (:1005), errors are logged and ``[]`` / ``False`` returned (:1028-1030) unless the
    store runs with ``GPU_STRICT``; ``add`` of an existing id overwrites the row (the HNSW intent,
    :370-375); ``remove`` tombstones the row so it can never be returned again.
    """

    def __init__(self, vector_dim: int, index_path: Path, config: Any = None, *, store=None, shard: int = 0):
        if store is None:
            raise ValueError("B200FlatIndex is created by wdbx_b200.VectorStore (it shares the store's device engine)")
        self.vector_dim = vector_dim
        self.index_path = index_path
        self.config = config or {}
        self._store = store
        self.shard = shard

    async def initialize(self):
        """Nothing to do: the engine is created with the store (indexing.py:840-843)."""

    async def shutdown(self):
        """Persistence is handled once by the store (indexing.py:845-849)."""

    # -- mutation ---------------------------------------------------------------------------
    def add(self, vector_id: str, vector: np.ndarray) -> bool:
        def run():
            return self._store._add_rows(self.shard, [vector_id], np.asarray(vector, dtype=np.float32)[None, :])
        return self._store._guard(False, run) is not False

    async def add_async(self, vector_id: str, vector: np.ndarray) -> bool:
        return await self._store._run(self.add, vector_id, vector)

    def batch_add(self, vectors: Dict[str, np.ndarray]) -> bool:
        if not vectors:
            return True
        def run():   # ragged input fails inside the guard: the operator boundary returns False, it never raises
            mat = np.stack([np.asarray(v, dtype=np.float32) for v in vectors.values()])
            return self._store._add_rows(self.shard, list(vectors.keys()), mat)
        return self._store._guard(False, run) is not False

    async def batch_add_async(self, vectors: Dict[str, np.ndarray]) -> bool:
        return await self._store._run(self.batch_add, vectors)

    def remove(self, vector_id: str) -> bool:
        return bool(self._store._guard(False, self._store._remove_row, vector_id, self.shard))

    async def remove_async(self, vector_id: str) -> bool:
        return await self._store._run(self.remove, vector_id)

    def clear(self) -> bool:
        return self._store._guard(False, self._store._clear_shard, self.shard) is not False

    async def clear_async(self) -> bool:
        return await self._store._run(self.clear)

    def optimize(self) -> bool:
        """A flat scan has nothing to optimise (indexing.py:1124-1147)."""
        return True

    async def optimize_async(self) -> bool:
        return True

    # -- search -----------------------------------------------------------------------------
    def search(self, query_vector: np.ndarray, limit: int = 10) -> List[Tuple[str, float]]:
        res = self._store._guard([], self._store._search_lists, np.asarray(query_vector, dtype=np.float32), limit,
                                 self.shard)
        return res[0] if res else []

    async def search_async(self, query_vector: np.ndarray, limit: int = 10) -> List[Tuple[str, float]]:
        return await self._store._run(self.search, query_vector, limit)

    # -- stats ------------------------------------------------------------------------------
    def size(self) -> int:
        return self._store._shard_live[self.shard]

    def get_stats(self) -> Dict[str, Any]:
        """Shape of FaissIndex.get_stats (indexing.py:1170-1183)."""
        return {
            "type": "b200_flat",
            "size": self.size(),
            "dimension": self.vector_dim,
            "gpu_enabled": True,
            "metric": self._store.metric,
            "dtype": self._store.dtype,
            "rows_stored": self._store._shard_count[self.shard],
        }
