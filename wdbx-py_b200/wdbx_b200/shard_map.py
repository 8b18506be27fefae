"""Shard placement: logical shard of an id, and where the n-th row of a shard lives.

Reference behaviour being replaced:
* ``VectorStore._get_shard_for_id`` (wdbx/core/vector_store.py:178-190): ``abs(hash(id)) % num_shards``.
  Python's ``str`` hash is salted per process, so the reference's placement is not reproducible
  across restarts -- and under one-process-per-GPU it would differ between ranks.  We use a
  deterministic CRC-32 instead (documented deviation; results do not depend on placement).
* ``ShardManager._allocate_shards`` (wdbx/core/distributed.py:547-654): shard -> node table,
  "fewest shards first".  Here every logical shard is striped row-wise over ALL ranks (row n of a
  shard lives on rank ``n % world`` at local row ``n // world``), which keeps every GPU within one
  row of perfectly balanced for any num_shards / world_size combination (the quick-start has
  num_shards=2 on up to 8 GPUs) and needs no communication to agree on.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Tuple


def stable_hash(vector_id: str) -> int:
    """Process-independent hash of an id (CRC-32 of its UTF-8 bytes; C speed)."""
    return zlib.crc32(vector_id.encode("utf-8"))


def shard_for_id(vector_id: str, num_shards: int) -> int:
    return stable_hash(vector_id) % num_shards


class ShardMap:
    def __init__(self, num_shards: int, world_size: int = 1):
        if num_shards < 1 or world_size < 1:
            raise ValueError("num_shards and world_size must be >= 1")
        self.num_shards = num_shards
        self.world_size = world_size

    def owner(self, n: int) -> Tuple[int, int]:
        """(rank, local_row) of the n-th row ever appended to a shard."""
        return n % self.world_size, n // self.world_size

    def local_count(self, total: int, rank: int) -> int:
        """How many of a shard's first `total` rows live on `rank`."""
        w = self.world_size
        return total // w + (1 if rank < total % w else 0)

    def allocation(self, shard_counts: List[int]) -> Dict[int, Dict]:
        """ShardManager.get_shard_info-style table (distributed.py:656-696) for stats."""
        return {
            s: {
                "shard": s,
                "partitions": [
                    {"rank": r, "rows": self.local_count(shard_counts[s], r)} for r in range(self.world_size)
                ],
            }
            for s in range(self.num_shards)
        }
