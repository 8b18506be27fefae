"""Device-resident VectorStore: the host-side mirror of wdbx/core/vector_store.py for the search path.

Same public surface as the reference class (``store / store_async / batch_store(_async) / search /
search_async / delete / get / count / clear / get_stats / update_metadata``; vector_store.py:22-815)
with the same argument meaning, result shape ``[(vector_id, score, metadata)]`` and error
behaviour, but:

* the per-shard index objects share ONE device engine (libwdbx_b200.so): an unfiltered
  ``search`` is a single kernel launch that returns the already merged global top-k instead of
  the reference's sequential per-shard loop + host sort (vector_store.py:323-330);
* with ``filter_metadata`` the engine returns one top-``limit`` list per shard, which is exactly
  the candidate set the reference builds before its post-filter (vector_store.py:323-342), so
  filtered results are identical to the reference's (including its truncation quirk);
* vectors live only in HBM (no host dict of ndarrays, vector_store.py:66): ``get`` reads the row
  back from the device;
* under ``torchrun`` (WORLD_SIZE > 1) the store is SPMD: every rank makes the same calls, rows of
  each shard are striped over the ranks (shard_map.py) and ``search`` is a collective
  (local top-k -> NCCL all-gather of packed keys -> merge kernel).

Additive API for the batch configs: ``search_batch`` and ``bulk_load``.
There is no CPU fallback: construction fails if the CUDA library or a device is missing.
"""
from __future__ import annotations

import asyncio
import bisect
import json
import logging
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .config import WDBXConfig
from .dist import DistContext
from .indexing import B200FlatIndex
from .shard_map import ShardMap, shard_for_id

logger = logging.getLogger(__name__)

ALL, EACH = _lib.ALL_SEGMENTS, _lib.EACH_SEGMENT


class _GidMap:
    """gid -> explicit string id, sparse.  Bulk-loaded rows have synthetic ids and no entry, so a
    10M-row bulk load does not create a 10M-element Python list (whose gen-2 GC traversal showed
    up as a 30 ms pause inside the search loop)."""

    __slots__ = ("_d", "_n")

    def __init__(self):
        self._d: Dict[int, str] = {}
        self._n = 0

    def __len__(self):
        return self._n

    def __getitem__(self, gid: int) -> Optional[str]:
        return self._d.get(gid)

    def __setitem__(self, gid: int, value: Optional[str]):
        if value is None:
            self._d.pop(gid, None)
        else:
            self._d[gid] = value

    def append(self, value: str):
        self._d[self._n] = value
        self._n += 1

    def skip(self, n: int):
        self._n += n

    def drop_all(self):
        self._d = {}

    def items(self):
        return sorted(self._d.items())

    @classmethod
    def restore(cls, n: int, pairs):
        m = cls()
        m._n = n
        m._d = {int(g): v for g, v in pairs}
        return m


class BatchResult:
    """Result of ``search_batch``: device-merged [B, k] arrays + lazy id mapping."""

    def __init__(self, gids: np.ndarray, scores: np.ndarray, counts: np.ndarray, store: "VectorStore"):
        self.gids, self.scores, self.counts, self._store = gids, scores, counts, store

    def ids(self) -> List[List[str]]:
        id_of = self._store._id_of       # one tolist() per array: numpy scalars cost more than the mapping itself
        return [[id_of(g) for g in row[:c]] for row, c in zip(self.gids.tolist(), self.counts.tolist())]

    def as_lists(self) -> List[List[Tuple[str, float]]]:
        id_of = self._store._id_of
        return [[(id_of(g), v) for g, v in zip(grow[:c], srow[:c])]
                for grow, srow, c in zip(self.gids.tolist(), self.scores.tolist(), self.counts.tolist())]


class VectorStore:
    def __init__(
        self,
        vector_dim: int,
        data_dir: Path,
        num_shards: int = 1,
        use_gpu: bool = True,
        index_type: str = "b200",
        config: Optional[WDBXConfig] = None,
        *,
        dist: Optional[DistContext] = None,
        _engine_factory=None,  # test seam (tests/ inject a numpy double); the product never sets it
    ):
        if not use_gpu:
            raise RuntimeError(
                "wdbx_b200.VectorStore implements only the enable_gpu=True path of WDBX; "
                "there is no CPU fallback (use the reference package for CPU search)")
        if index_type not in ("b200", "flat", "faiss", "hnsw"):
            raise ValueError(f"Unsupported index type: {index_type}")
        if num_shards < 1 or num_shards > _lib.MAX_SEGMENTS:
            raise ValueError(f"num_shards must be in [1, {_lib.MAX_SEGMENTS}]")
        self.vector_dim = int(vector_dim)
        self._qstruct = struct.Struct(f"{self.vector_dim}f")   # list-of-floats query -> fp32 bytes
        self.data_dir = Path(data_dir)
        self.num_shards = int(num_shards)
        self.use_gpu = True
        self.index_type = "b200"
        self.config = config if config is not None else WDBXConfig({})
        self.metric = str(self.config.get("GPU_METRIC", "cosine")).lower()
        if self.metric not in _lib.METRICS:
            raise ValueError(f"unknown GPU_METRIC {self.metric!r}")
        self.dtype = str(self.config.get("GPU_DTYPE", "fp32")).lower()
        self.strict = bool(self.config.get("GPU_STRICT", False))
        self.prefilter = bool(self.config.get("GPU_PREFILTER", False))   # opt-in: changes filtered results
        self._allow_cache: Dict[str, Any] = {}
        self._version = 0                                                # bumped by every mutation

        self.dist = dist if dist is not None else DistContext.from_env(self.config.get("GPU_DEVICE", None))
        self.shard_map = ShardMap(self.num_shards, self.dist.world)

        # host-side bookkeeping (ids and metadata only; vectors live in HBM)
        self.metadata: Dict[str, Dict[str, Any]] = {}
        self._loc: Dict[str, Tuple[int, int, int]] = {}      # id -> (shard, n-th row of the shard, gid)
        self._gid_to_id = _GidMap()
        self._bulk: List[Tuple[int, int, str]] = []          # (gid0, gid1, prefix) ranges from bulk_load
        self._bulk_starts: List[int] = []
        self._bulk_rows: Dict[int, Tuple[int, int]] = {}     # gid0 -> (first n per shard base, n rows) for lookups
        self._bulk_dead: set = set()
        self._bulk_cleared: Dict[int, set] = {}              # gid0 -> shards cleared since that bulk_load
        self._row_gids: List[List[np.ndarray]] = [[] for _ in range(self.num_shards)]  # gid of every row, per shard
        self._shard_count = [0] * self.num_shards            # rows ever appended per shard (all ranks)
        self._shard_live = [0] * self.num_shards
        self._lock = threading.RLock()
        self.thread_pool = ThreadPoolExecutor(
            max_workers=int(self.config.get("VECTOR_STORE_THREADS", min(8, os.cpu_count() or 4))))

        self._create_dirs()
        # GPU_DEVICES: one ordinary process drives several GPUs (multi_engine.py); under torchrun every rank has one
        from .multi_engine import MultiEngine, parse_devices

        devices = parse_devices(self.config.get("GPU_DEVICES", None)) if self.dist.world == 1 else None
        self.devices = devices if devices and len(devices) > 1 else [self.dist.device if not devices else devices[0]]
        if len(self.devices) > 1:
            self.engine = MultiEngine(self.devices, self.vector_dim, self.dtype, self.num_shards,
                                      _engine_factory=_engine_factory,
                                      _group_factory=getattr(_engine_factory, "group_factory", None))
        elif _engine_factory is not None:
            self.engine = _engine_factory(self.devices[0], self.vector_dim, self.dtype, self.num_shards)
        else:
            from .engine import Engine  # raises ImportError / B200Error loudly when the GPU path is unusable

            self.engine = Engine(self.devices[0], self.vector_dim, self.dtype, self.num_shards)
        if bool(self.config.get("GPU_OVERLAP", False)) and hasattr(self.engine, "set_option"):
            self.engine.set_option("overlap", 1)
        cap = int(self.config.get("GPU_CAPACITY_ROWS", 0) or 0)
        if cap > 0:
            for s in range(self.num_shards):
                self.engine.reserve(s, self.shard_map.local_count(cap, self.dist.rank))
        # one box, several GPUs: fuse the cross-GPU merge into the scan kernel (NVLink P2P key push)
        self._fused = False
        if (self.dist.world > 1 and _engine_factory is None and hasattr(self.engine, "exchange_setup")
                and bool(self.config.get("GPU_FUSED_EXCHANGE", True))):
            try:
                self.engine.exchange_setup(self.dist.rank, self.dist.world, self.dist.all_gather_bytes)
                self._fused = True
            except Exception as e:  # e.g. > 8 ranks or no peer access: NCCL all-gather + merge kernel instead
                logger.warning(f"fused exchange unavailable, using NCCL all-gather: {e}")
        # async front-end: coalesce concurrent single-query requests (single process only: under SPMD
        # every rank would have to form identical batches)
        self._batcher = None
        bmax = int(self.config.get("GPU_BATCH_MAX", 64) or 0)
        if self.dist.world == 1 and bmax > 1:
            from .batcher import MicroBatcher

            self._batcher = MicroBatcher(self, bmax, float(self.config.get("GPU_BATCH_WINDOW_US", 200)))
        self._init_indices()
        self._load_data()
        logger.info("VectorStore initialized: %d shards on %d GPU(s), dim=%d, metric=%s, dtype=%s",
                    self.num_shards, max(self.dist.world, len(self.devices)), self.vector_dim, self.metric, self.dtype)

    # ------------------------------------------------------------------ setup
    def _create_dirs(self):
        """Same on-disk layout as the reference (vector_store.py:88-109)."""
        for sub in ("vectors", "metadata", "indices"):
            (self.data_dir / sub).mkdir(parents=True, exist_ok=True)
        for shard in range(self.num_shards):
            (self.data_dir / f"shard_{shard}").mkdir(parents=True, exist_ok=True)

    def _init_indices(self):
        self.indices = [
            B200FlatIndex(self.vector_dim, self.data_dir / f"shard_{s}" / "index", self.config, store=self, shard=s)
            for s in range(self.num_shards)
        ]

    def _get_shard_for_id(self, vector_id: str) -> int:
        return shard_for_id(vector_id, self.num_shards)

    async def initialize(self):
        await asyncio.gather(*[ix.initialize() for ix in self.indices])

    async def shutdown(self):
        """Reference: save metadata + vectors, then every index (vector_store.py:202-217)."""
        loop = asyncio.get_event_loop()
        await loop.run_in_executor(self.thread_pool, self.save)
        self.thread_pool.shutdown(wait=True)
        self.close()

    def close(self):
        if getattr(self, "_batcher", None) is not None:
            self._batcher.close()
            self._batcher = None
        eng = getattr(self, "engine", None)
        if eng is not None:
            eng.close()
            self.engine = None

    def _save_metadata(self):
        if self.dist.rank != 0:
            return
        try:
            with open(self.data_dir / "metadata" / "metadata.json", "w") as f:
                json.dump(self.metadata, f)
        except Exception as e:  # reference logs and continues (vector_store.py:165-166)
            logger.error(f"Error saving metadata: {e}")

    # ------------------------------------------------------------------ persistence
    # Reference: metadata.json + a pickle of the id -> ndarray dict (vector_store.py:158-176) and one
    # index file + mapping pickle per shard (indexing.py:805-838).  Here: metadata.json unchanged,
    # ids / placement in vectors/state.json, and every rank dumps its device partition of every
    # shard as a raw .npy that is appended straight back to HBM at start-up.
    STATE_VERSION = 1

    def save(self) -> bool:
        try:
            with self._lock:
                self._save_metadata()
                rank, world = self.dist.rank, self.dist.world
                for s in range(self.num_shards):
                    n_local = self.shard_map.local_count(self._shard_count[s], rank)
                    rows = self.engine.read_rows(s, 0, n_local)
                    np.save(self.data_dir / f"shard_{s}" / f"rows.rank{rank}of{world}.npy", rows)
                if rank == 0:
                    for s in range(self.num_shards):
                        order = (np.concatenate(self._row_gids[s]) if self._row_gids[s]
                                 else np.empty(0, np.uint32))
                        np.save(self.data_dir / f"shard_{s}" / "row_gids.npy", order.astype(np.uint32))
                    state = {
                        "version": self.STATE_VERSION, "dim": self.vector_dim, "dtype": self.dtype,
                        "num_shards": self.num_shards, "world": world, "shard_count": self._shard_count,
                        "next_gid": len(self._gid_to_id),
                        "ids": [[g, v] for g, v in self._gid_to_id.items()],
                        "bulk": [[g0, g1, p, list(self._bulk_rows[g0])] for g0, g1, p in self._bulk],
                        "bulk_dead": sorted(self._bulk_dead),
                        "bulk_cleared": [[g0, sorted(v)] for g0, v in self._bulk_cleared.items()],
                    }
                    tmp = self.data_dir / "vectors" / "state.json.tmp"
                    tmp.write_text(json.dumps(state))
                    tmp.replace(self.data_dir / "vectors" / "state.json")
            self.dist.barrier()
            if rank == 0:   # partitions of an earlier run with another rank count are superseded now
                for s in range(self.num_shards):
                    for f in (self.data_dir / f"shard_{s}").glob("rows.rank*of*.npy"):
                        if not f.name.endswith(f"of{world}.npy"):
                            f.unlink(missing_ok=True)
            return True
        except Exception as e:
            logger.error(f"Error saving vectors: {e}")
            if self.strict:
                raise
            return False

    def _load_partition(self, shard: int, count: int, rank: int, world: int, saved_world: int) -> np.ndarray:
        """This rank's rows of a shard, in local order.  Position n of the shard lives on rank n % world at
        local row n // world; a store saved by a different number of ranks is RE-STRIPED: position n is
        fetched from the saved file of rank n % saved_world, row n // saved_world (memory-mapped, so a rank
        only touches the rows it keeps)."""
        d = self.data_dir / f"shard_{shard}"
        if saved_world == world:
            return np.load(d / f"rows.rank{rank}of{world}.npy")
        pos = np.arange(rank, count, world, dtype=np.int64)
        rows = np.empty((pos.shape[0], self.vector_dim), dtype=np.float32)
        src_rank, src_row = pos % saved_world, pos // saved_world
        for r in range(saved_world):
            m = src_rank == r
            if m.any():
                part = np.load(d / f"rows.rank{r}of{saved_world}.npy", mmap_mode="r")
                rows[m] = part[src_row[m]]
        return rows

    def _load_data(self) -> int:
        """Start-up load (reference: vector_store.py:136-156): rebuild the device partitions from disk."""
        path = self.data_dir / "vectors" / "state.json"
        if not path.exists():
            return 0
        try:
            state = json.loads(path.read_text())
            rank, world = self.dist.rank, self.dist.world
            for key, have in (("version", self.STATE_VERSION), ("dim", self.vector_dim), ("dtype", self.dtype),
                              ("num_shards", self.num_shards)):
                if state.get(key) != have:
                    raise ValueError(f"saved store has {key}={state.get(key)!r}, this instance {have!r}")
            saved_world = int(state.get("world", world))   # may differ: partitions are re-striped below
            n_gid = int(state["next_gid"])
            gid_to_id = _GidMap.restore(n_gid, state["ids"])
            bulk = [(g0, g1, p) for g0, g1, p, _ in state["bulk"]]
            dead = np.ones(n_gid, bool)           # explicit rows without an id were deleted
            if state["ids"]:
                dead[np.asarray([g for g, _ in state["ids"]], np.int64)] = False
            for g0, g1, _ in bulk:
                dead[g0:g1] = False
            dead[np.asarray(state["bulk_dead"], np.int64)] = True
            loc: Dict[str, Tuple[int, int, int]] = {}
            live = [0] * self.num_shards
            row_gids: List[List[np.ndarray]] = [[] for _ in range(self.num_shards)]
            for s in range(self.num_shards):
                order = np.load(self.data_dir / f"shard_{s}" / "row_gids.npy")
                if order.shape[0] != state["shard_count"][s]:
                    raise ValueError(f"shard {s}: row_gids.npy does not match state.json")
                rows = self._load_partition(s, order.shape[0], rank, world, saved_world)
                mine = order[rank::world]
                if rows.shape != (mine.shape[0], self.vector_dim):
                    raise ValueError(f"shard {s}: partition file has shape {rows.shape}")
                if mine.shape[0]:
                    self.engine.append(s, rows, gids=mine)
                    for local in np.flatnonzero(dead[mine]):
                        self.engine.tombstone(s, int(local), True)
                for n, g in enumerate(order.tolist()):
                    v = gid_to_id[g]
                    if v is not None:
                        loc[v] = (s, n, g)
                live[s] = int((~dead[order]).sum())
                row_gids[s] = [order]
            self._gid_to_id, self._loc, self._row_gids = gid_to_id, loc, row_gids
            self._bulk = bulk
            self._bulk_starts = [g0 for g0, _, _ in bulk]
            self._bulk_rows = {g0: tuple(base) for g0, _, _, base in state["bulk"]}
            self._bulk_dead = set(state["bulk_dead"])
            self._bulk_cleared = {int(g0): set(v) for g0, v in state.get("bulk_cleared", [])}
            self._shard_count = list(state["shard_count"])
            self._shard_live = live
            self._version += 1
            meta = self.data_dir / "metadata" / "metadata.json"
            if meta.exists():
                self.metadata = json.loads(meta.read_text())
            logger.info(f"Loaded {self.count()} vectors from {self.data_dir}")
            return self.count()
        except Exception as e:
            logger.error(f"Error loading vectors: {e}")
            try:
                self.engine.clear(ALL)
            except Exception:
                pass
            if self.strict:
                raise
            return 0

    # ------------------------------------------------------------------ error convention
    def _guard(self, default, fn, *args):
        """Reference convention: index errors -- a query or vector of the wrong dimension included -- are logged and
        swallowed (indexing.py:903-905, :1028-1030: the operator boundary never raises); GPU_STRICT re-raises.  The
        facade validates dimensions itself and raises ValueError as the reference's does (wdbx.py:323-326)."""
        try:
            return fn(*args)
        except Exception as e:
            if self.strict:
                raise
            logger.error(f"Error in B200 index operation {getattr(fn, '__name__', fn)}: {e}")
            return default

    async def _run(self, fn, *args):
        loop = asyncio.get_event_loop()
        return await loop.run_in_executor(self.thread_pool, fn, *args)

    # ------------------------------------------------------------------ id bookkeeping
    def _id_of(self, gid: int) -> str:
        """gid -> id; on the result path of every search (k calls per query: kept free of Python-level method calls)"""
        v = self._gid_to_id._d.get(gid)
        if v is not None:
            return v
        starts = self._bulk_starts
        if starts:
            i = bisect.bisect_right(starts, gid) - 1
            if i >= 0:
                g0, g1, prefix = self._bulk[i]
                if gid < g1:
                    return f"{prefix}{gid - g0}"
        return str(gid)  # same fallback as the reference's index_to_id.get(idx, str(idx)) (indexing.py:1021)

    def _locate(self, vector_id: str) -> Optional[Tuple[int, int, int]]:
        loc = self._loc.get(vector_id)
        if loc is not None:
            return loc
        for g0, g1, prefix in self._bulk:
            if vector_id.startswith(prefix):
                tail = vector_id[len(prefix):]
                # canonical decimal only: "v01" or a unicode digit must not alias row 1
                if tail.isascii() and tail.isdigit() and (tail == "0" or tail[0] != "0"):
                    gid = g0 + int(tail)
                    if gid < g1 and gid not in self._bulk_dead:
                        i = gid - g0
                        base = self._bulk_rows[g0]
                        shard = i % self.num_shards
                        if shard in self._bulk_cleared.get(g0, ()):
                            return None
                        return shard, base[shard] + i // self.num_shards, gid
        return None

    # ------------------------------------------------------------------ mutation
    def _add_rows(self, shard: int, ids: Sequence[str], mat: np.ndarray) -> bool:
        """Append (or overwrite, for ids already present) rows of one shard."""
        if mat.ndim != 2 or mat.shape[1] != self.vector_dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.vector_dim}, got {mat.shape[-1]}")
        rank, world = self.dist.rank, self.dist.world
        with self._lock:
            self._version += 1       # every mutation path ends here or in _remove_row / _clear_shard: cached allow bitmaps die
            fresh: List[int] = []
            for i, vid in enumerate(ids):
                loc = self._locate(vid)
                if loc is None:
                    fresh.append(i)
                    continue
                s0, n0, _gid = loc
                owner, local = self.shard_map.owner(n0)
                if owner == rank:
                    self.engine.overwrite(s0, local, mat[i])
            if not fresh:
                return True
            n0 = self._shard_count[shard]
            g0 = len(self._gid_to_id)
            m = len(fresh)
            ns = np.arange(n0, n0 + m)
            gids = np.arange(g0, g0 + m, dtype=np.uint32)
            mine = (ns % world) == rank
            # gid -> id FIRST: searches do not take the store lock (the reference's thread pool runs them next to
            # `add`, indexing.py:692), and one that lands between the device append and the bookkeeping below must
            # already resolve the new rows' ids instead of the str(gid) fallback
            for i in fresh:
                self._gid_to_id.append(ids[i])
            try:
                if mine.any():
                    sel = np.asarray(fresh)[mine]
                    rows = mat if len(sel) == mat.shape[0] else mat[sel]
                    first = self.engine.append(shard, rows, gids=gids[mine])
                    expect = self.shard_map.owner(int(ns[mine][0]))[1]
                    if first != expect:
                        raise RuntimeError(f"shard {shard}: device row {first} != expected {expect} (store out of sync)")
            except BaseException:
                for g in range(g0, g0 + m):     # the gids stay consumed (they only have to be unique)
                    self._gid_to_id[g] = None
                raise
            for j, i in enumerate(fresh):
                self._loc[ids[i]] = (shard, n0 + j, g0 + j)
            self._row_gids[shard].append(gids)
            self._shard_count[shard] += m
            self._shard_live[shard] += m
        return True

    def _remove_row(self, vector_id: str, shard: Optional[int] = None) -> bool:
        with self._lock:
            loc = self._locate(vector_id)
            if loc is None or (shard is not None and loc[0] != shard):
                return False
            s, n, gid = loc
            owner, local = self.shard_map.owner(n)
            if owner == self.dist.rank:
                self.engine.tombstone(s, local, True)
            if vector_id in self._loc:
                del self._loc[vector_id]
                self._gid_to_id[gid] = None
            else:
                self._bulk_dead.add(gid)
            # an id that is gone has no metadata -- also when the row was removed through the shard's index facade
            # (FaissIndex.remove knows nothing of the store's metadata, indexing.py:1050-1074; a stale entry would
            # resurface under a later row of the same id)
            self.metadata.pop(vector_id, None)
            self._shard_live[s] -= 1
            self._version += 1
            return True

    def _clear_shard(self, shard: int) -> bool:
        with self._lock:
            self.engine.clear(shard)
            for vid in [v for v, loc in self._loc.items() if loc[0] == shard]:
                gid = self._loc.pop(vid)[2]
                self._gid_to_id[gid] = None
                self.metadata.pop(vid, None)
            # bulk rows of this shard are gone too: their ids must stop resolving to (reused) row positions
            for g0, g1, prefix in self._bulk:
                self._bulk_cleared.setdefault(g0, set()).add(shard)
                if self.metadata:
                    for i in range(shard, g1 - g0, self.num_shards):
                        vid = f"{prefix}{i}"
                        if vid not in self._loc:   # (re-stored since as an explicit row of another shard: keeps its metadata)
                            self.metadata.pop(vid, None)
            self._row_gids[shard] = []
            self._version += 1
            self._shard_count[shard] = 0
            self._shard_live[shard] = 0
        return True

    def store(self, vector_id: str, vector: List[float], metadata: Optional[Dict[str, Any]] = None) -> bool:
        """Reference: vector_store.py:219-256 (returns False on error, never raises)."""
        try:
            vec = np.array(vector, dtype=np.float32)
            shard = self._get_shard_for_id(vector_id)
            with self._lock:
                loc = self._locate(vector_id)
                # metadata BEFORE the row becomes searchable, as the reference does (vector_store.py:241-246): a
                # concurrent search that finds the new row must see (and post-filter on) its metadata
                had, prev = vector_id in self.metadata, self.metadata.get(vector_id)
                self.metadata[vector_id] = metadata or {}
                ok = False
                try:
                    ok = self.indices[loc[0] if loc else shard].add(vector_id, vec)
                finally:
                    if ok:
                        self._version += 1
                    elif had:
                        self.metadata[vector_id] = prev
                    else:
                        self.metadata.pop(vector_id, None)
            if ok and self.config.get("VECTOR_STORE_SAVE_IMMEDIATELY", False):
                self._save_metadata()
            return bool(ok)
        except Exception as e:
            logger.error(f"Error storing vector: {e}")
            return False

    async def store_async(self, vector_id: str, vector: List[float],
                          metadata: Optional[Dict[str, Any]] = None) -> bool:
        """Reference: vector_store.py:258-299."""
        return await self._run(self.store, vector_id, vector, metadata)

    def batch_store(self, vectors: Dict[str, List[float]],
                    metadata: Optional[Dict[str, Dict[str, Any]]] = None) -> int:
        """Reference: vector_store.py:720-766 -- group by shard, one batch_add per shard."""
        metadata = metadata or {}
        shard_vectors: Dict[int, Dict[str, np.ndarray]] = {}
        with self._lock:
            for vector_id, vector in vectors.items():
                try:
                    vec = np.array(vector, dtype=np.float32)
                    if vec.shape != (self.vector_dim,):
                        raise ValueError(f"Vector dimension mismatch: expected {self.vector_dim}, got {vec.shape}")
                except Exception as e:   # one bad vector does not take its shard's batch down; the count says so
                    if self.strict:
                        raise
                    logger.error(f"Error storing vector {vector_id!r}: {e}")
                    continue
                loc = self._locate(vector_id)
                shard = loc[0] if loc else self._get_shard_for_id(vector_id)
                shard_vectors.setdefault(shard, {})[vector_id] = vec
            stored = 0
            for shard, vecs in shard_vectors.items():
                prev = {vid: self.metadata[vid] for vid in vecs if vid in self.metadata}
                for vid in vecs:                      # metadata first (see store)
                    self.metadata[vid] = metadata.get(vid, {})
                ok = False
                try:
                    ok = self.indices[shard].batch_add(vecs)
                finally:
                    if ok:
                        stored += len(vecs)
                        self._version += 1
                    else:
                        for vid in vecs:
                            if vid in prev:
                                self.metadata[vid] = prev[vid]
                            else:
                                self.metadata.pop(vid, None)
        if self.config.get("VECTOR_STORE_SAVE_IMMEDIATELY", False):
            self._save_metadata()
        return stored

    async def batch_store_async(self, vectors: Dict[str, List[float]],
                                metadata: Optional[Dict[str, Dict[str, Any]]] = None) -> int:
        return await self._run(self.batch_store, vectors, metadata)

    def bulk_load(self, rows, id_prefix: str = "v") -> int:
        """Additive bulk ingest: append an [n, dim] fp32 matrix (numpy, or
        a CUDA tensor holding THIS rank's rows when ``rows`` is a dict {"local": tensor, "total": n})
        with synthetic ids ``f"{id_prefix}{i}"``.  Row i goes to shard ``i % num_shards``; inside a shard
        rows are striped over ranks as usual.  No per-row Python work."""
        S, world, rank = self.num_shards, self.dist.world, self.dist.rank
        with self._lock:
            if isinstance(rows, dict):
                if S != 1:
                    raise ValueError("pre-partitioned bulk_load needs num_shards == 1")
                n = int(rows["total"])
                local = rows["local"]
            else:
                n = int(rows.shape[0])
                local = None
            for g0_, g1_, p_ in self._bulk:
                if p_ == id_prefix:
                    raise ValueError(f"id_prefix {id_prefix!r} already used by a bulk_load")
                # "q1" + "23" and "q12" + "3" would both be the id "q123": a prefix that extends (or is extended by)
                # another bulk's prefix with digits only is ambiguous
                longer, shorter = (id_prefix, p_) if len(id_prefix) > len(p_) else (p_, id_prefix)
                rest = longer[len(shorter):]
                if longer.startswith(shorter) and rest.isascii() and rest.isdigit():
                    raise ValueError(f"id_prefix {id_prefix!r} is ambiguous next to the bulk prefix {p_!r}")
            for vid in self._loc:   # an explicit id "v5" would shadow bulk row 5 of prefix "v"
                if vid.startswith(id_prefix):
                    tail = vid[len(id_prefix):]
                    if tail.isascii() and tail.isdigit() and (tail == "0" or tail[0] != "0") and int(tail) < n:
                        raise ValueError(f"id_prefix {id_prefix!r} collides with the existing id {vid!r}")
            g0 = len(self._gid_to_id)
            base = list(self._shard_count)
            # the synthetic id range FIRST (see _add_rows: concurrent searches must resolve the new rows' ids)
            self._gid_to_id.skip(n)
            self._bulk.append((g0, g0 + n, id_prefix))
            self._bulk_starts.append(g0)
            self._bulk_rows[g0] = tuple(base)
            s = 0
            try:
                for s in range(S):
                    cnt = (n - s + S - 1) // S if n > s else 0
                    if cnt == 0:
                        continue
                    n0 = base[s]
                    ns = np.arange(n0, n0 + cnt)
                    mine = (ns % world) == rank
                    src_idx = s + S * np.arange(cnt)
                    gids = (g0 + src_idx).astype(np.uint32)
                    if local is not None:
                        if int(mine.sum()) != int(local.shape[0]):
                            raise ValueError("local tensor does not match this rank's stripe")
                        self.engine.append(s, local, gids=gids[mine])
                    elif mine.any():
                        sel = src_idx[mine]
                        part = rows[sel] if not (S == 1 and world == 1) else rows
                        self.engine.append(s, part, gids=gids[mine])
                    self._row_gids[s].append(gids)
                    self._shard_count[s] += cnt
                    self._shard_live[s] += cnt
            except BaseException:
                # shards s .. S-1 never received their rows: their ids of this bulk must not resolve to positions
                self._bulk_cleared.setdefault(g0, set()).update(range(s, S))
                self._version += 1
                raise
            self._version += 1
        return n

    def delete(self, vector_id: str) -> bool:
        """Reference: vector_store.py:465-492."""
        with self._lock:
            if self._locate(vector_id) is None:
                return False
            ok = self._guard(False, self._remove_row, vector_id)   # (drops the metadata with the row; a failed removal keeps both)
        if self.config.get("VECTOR_STORE_SAVE_IMMEDIATELY", False):
            self._save_metadata()
        return bool(ok)

    async def delete_async(self, vector_id: str) -> bool:
        return await self._run(self.delete, vector_id)

    def update_metadata(self, vector_id: str, metadata: Dict[str, Any]) -> bool:
        """Reference: vector_store.py:526-548."""
        with self._lock:
            if self._locate(vector_id) is None:
                return False
            self.metadata[vector_id] = metadata
            self._version += 1
        if self.config.get("VECTOR_STORE_SAVE_IMMEDIATELY", False):
            self._save_metadata()
        return True

    async def update_metadata_async(self, vector_id: str, metadata: Dict[str, Any]) -> bool:
        return self.update_metadata(vector_id, metadata)

    def get(self, vector_id: str) -> Optional[Tuple[List[float], Dict[str, Any]]]:
        """Reference: vector_store.py:579-596; the vector is read back from HBM."""
        with self._lock:
            loc = self._locate(vector_id)
            if loc is None:
                return None
            s, n, _gid = loc
            owner, local = self.shard_map.owner(n)
            if self.dist.world == 1:
                vec = self.engine.read_row(s, local)
            else:
                vec = self.engine.read_row(s, local) if owner == self.dist.rank else np.zeros(self.vector_dim, np.float32)
                vec = self.dist.broadcast_array(vec, owner)
            return vec.tolist(), self.metadata.get(vector_id, {})

    async def get_async(self, vector_id: str):
        return self.get(vector_id)

    def count(self) -> int:
        return sum(self._shard_live)

    def clear(self) -> int:
        """Reference: vector_store.py:620-641."""
        with self._lock:
            count = self.count()
            self.engine.clear(ALL)
            self.metadata = {}
            self._loc = {}
            self._gid_to_id.drop_all()  # gids keep growing: keys stay unique
            self._bulk, self._bulk_starts, self._bulk_rows, self._bulk_dead = [], [], {}, set()
            self._bulk_cleared = {}
            self._shard_count = [0] * self.num_shards
            self._shard_live = [0] * self.num_shards
            self._row_gids = [[] for _ in range(self.num_shards)]
            self._version += 1
        self._save_metadata()
        return count

    async def clear_async(self) -> int:
        return await self._run(self.clear)

    def optimize(self) -> bool:
        return True

    async def optimize_async(self) -> bool:
        return True

    # ------------------------------------------------------------------ search
    def _search_arrays(self, Q: np.ndarray, k: int, sel: int, metric: Optional[str] = None):
        """(scores, gids, counts) numpy arrays; leading dim = num_shards when sel == EACH."""
        metric = metric or self.metric
        if self.dist.world == 1:
            return self.engine.search_host(Q, k, metric=metric, per_segment=(sel == EACH),
                                           segment=(sel if sel >= 0 else ALL))
        # SPMD: local top-k on every rank -> on-device NVLink exchange, or all-gather of packed keys + merge kernel
        if sel == ALL and self._fused and Q.shape[0] <= self.engine.XCHG_MAX_B and k <= self.engine.XCHG_MAX_K:
            scores, gids, counts = self.engine.search_exchange_host(Q, k, metric)   # one C call: H2D, kernels, D2H
            if (counts < 0).any():
                raise RuntimeError("fused exchange timed out: a peer rank did not join the collective search")
            return scores, gids, counts
        qd = self.engine.upload(Q)
        if sel == EACH:
            import torch

            keys = torch.cat([self.engine.search(qd, k, metric, segment=s)["keys"] for s in range(self.num_shards)])
        else:
            keys = self.engine.search(qd, k, metric, segment=sel)["keys"]
        merged = self.engine.merge(self.dist.all_gather_keys(keys))
        scores, gids, counts = self._unpack(merged, keys.shape[0], k)
        if sel == EACH:
            B = Q.shape[0]
            return (scores.reshape(self.num_shards, B, k), gids.reshape(self.num_shards, B, k),
                    counts.reshape(self.num_shards, B))
        return scores, gids, counts

    @staticmethod
    def _unpack(out, B: int, k: int):
        """Device result dict -> (scores, gids, counts) numpy; one D2H when the engine packed it."""
        if "packed" in out:
            from .engine import unpack_out

            return unpack_out(out, B, k)
        return out["scores"].cpu().numpy(), out["gids"].cpu().numpy(), out["counts"].cpu().numpy()

    def search_device(self, q_dev, limit: int = 10, metric: Optional[str] = None):
        """Additive, fully device-resident search: ``q_dev`` is a CUDA fp32 tensor [B, dim] on this
        rank's GPU (the same queries on every rank); returns CUDA tensors ``keys / scores / gids /
        counts`` of the global top-``limit``.  No host copy, no synchronisation."""
        metric = metric or self.metric
        if self.dist.world == 1:
            return self.engine.search(q_dev, limit, metric)
        if self._fused and q_dev.shape[0] <= self.engine.XCHG_MAX_B and limit <= self.engine.XCHG_MAX_K:
            return self.engine.search_exchange(q_dev, limit, metric)
        out = self.engine.search(q_dev, limit, metric)
        return self.engine.merge(self.dist.all_gather_keys(out["keys"]))

    def _search_lists(self, q: np.ndarray, limit: int, sel: int) -> List[List[Tuple[str, float]]]:
        """Best-first [(id, score)] lists for ONE query: one list (merged / single shard) or one per shard."""
        q = np.asarray(q, dtype=np.float32).reshape(-1)
        if q.shape[0] != self.vector_dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.vector_dim}, got {q.shape[0]}")
        nlists = self.num_shards if sel == EACH else 1
        if sel >= 0:
            scope = self._shard_live[sel]
        elif sel == EACH:
            scope = max(self._shard_live)
        else:
            scope = sum(self._shard_live)
        k = min(int(limit), scope, _lib.MAX_K)   # limit clipped to the live size (indexing.py:1005)
        if k <= 0:
            return [[] for _ in range(nlists)]
        if limit > _lib.MAX_K and scope > _lib.MAX_K:
            if sel == ALL and self.dist.world == 1 and hasattr(self.engine, "search_filtered_host"):
                return [self._search_paged(q, min(int(limit), scope))]
            logger.warning("limit=%d clipped to the engine maximum of %d", limit, _lib.MAX_K)
        scores, gids, counts = self._search_arrays(q[None, :], k, sel)
        id_of = self._id_of
        if sel == EACH:   # .tolist(): one conversion per array instead of one numpy scalar per element
            return [[(id_of(g), s) for g, s in zip(gids[i, 0, : counts[i, 0]].tolist(), scores[i, 0, : counts[i, 0]].tolist())]
                    for i in range(nlists)]
        c = int(counts[0])
        return [[(id_of(g), s) for g, s in zip(gids[0, :c].tolist(), scores[0, :c].tolist())]]

    def _search_paged(self, q: np.ndarray, want: int) -> List[Tuple[str, float]]:
        """More than MAX_K neighbours (the reference returns up to `limit` rows, indexing.py:1005): successive exact passes
        of MAX_K, each over the rows NOT returned so far -- they are masked out with the per-row allow bitmaps of the
        pre-filter path (the select kernel consults them), so page n+1 continues exactly where page n stopped: the
        concatenation is the exact top-`want`, ties included (the key order is total).  One full scan per page;
        single process (one engine or a device group), unfiltered searches."""
        out: List[Tuple[str, float]] = []
        id_of = self._id_of
        with self._lock:     # the bitmaps cover exactly the rows present now (see _search_prefiltered)
            orders = [np.concatenate(r) if r else np.empty(0, np.uint32) for r in self._row_gids]   # ascending gids per shard
            maps = []
            for o in orders:
                bits = np.ones(o.shape[0], dtype=np.uint8)
                pad = (-bits.shape[0]) % 32
                maps.append(np.packbits(np.concatenate([bits, np.zeros(pad, np.uint8)]), bitorder="little").view(np.uint32)
                            if bits.size else None)      # (None = "all rows" of an empty shard: nothing to copy)
            while len(out) < want:
                k = min(_lib.MAX_K, want - len(out))
                scores, gids, counts = self.engine.search_filtered_host(q[None, :], k, self.metric, float("-inf"), maps)
                c = int(counts[0])
                if c <= 0:
                    break
                page = gids[0, :c]
                for s_, o in enumerate(orders):          # clear the bits of this page's rows in their shard
                    if o.shape[0] == 0:
                        continue
                    pos = np.searchsorted(o, page)
                    hit = (pos < o.shape[0]) & (o[np.minimum(pos, o.shape[0] - 1)] == page)
                    p = pos[hit]
                    np.bitwise_and.at(maps[s_], p >> 5, ~(np.uint32(1) << (p & 31).astype(np.uint32)))
                out.extend((id_of(g), sc) for g, sc in zip(page.tolist(), scores[0, :c].tolist()))
                if c < k:
                    break
        return out

    def _query_array(self, query_vector) -> np.ndarray:
        """fp32 array of a query given as a list of Python floats (the reference API's form).  struct.pack does
        the same double -> float rounding as numpy in a third of the time (13 vs 33 us for 768 floats, which is
        visible next to a 370 us search on 8 GPUs); anything it rejects goes through numpy as before."""
        if type(query_vector) is list and len(query_vector) == self.vector_dim:
            try:
                return np.frombuffer(bytearray(self._qstruct.pack(*query_vector)), dtype=np.float32)   # writable
            except (struct.error, OverflowError, TypeError):
                pass
        return np.array(query_vector, dtype=np.float32)

    def search(self, query_vector: List[float], limit: int = 10, threshold: float = 0.0,
               filter_metadata: Optional[Dict[str, Any]] = None) -> List[Tuple[str, float, Dict[str, Any]]]:
        """Reference: vector_store.py:301-353 (same result list, same filter / threshold order)."""
        query_np = self._query_array(query_vector)
        if filter_metadata and self.prefilter and (self.dist.world == 1 or self._fused):
            # opt-in (GPU_PREFILTER): exact top-`limit` AMONG the matching rows + threshold push-down
            lists = self._guard([], self._search_prefiltered, query_np, limit, threshold, filter_metadata)
            return [(vid, score, self.metadata.get(vid, {})) for vid, score in (lists[0] if lists else [])]
        sel = EACH if filter_metadata else ALL
        lists = self._guard([], self._search_lists, query_np, limit, sel)
        all_results: List[Tuple[str, float]] = []
        for results in lists:
            all_results.extend(results)
        if len(lists) > 1:
            all_results.sort(key=lambda x: x[1], reverse=True)
        if threshold > 0:
            all_results = [r for r in all_results if r[1] >= threshold]
        if filter_metadata:
            all_results = [r for r in all_results if self._matches_filter(r[0], filter_metadata)]
        all_results = all_results[:limit]
        return [(vid, score, self.metadata.get(vid, {})) for vid, score in all_results]

    # ------------------------------------------------------------------ opt-in device-side pre-filter
    def _allow_bitmaps(self, filter_metadata: Dict[str, Any]):
        """Per-shard uint32 bitmaps over the rows of this rank's partitions: bit = metadata matches.
        Cached per (filter, store version); evaluating the Mongo-style filter is host work over the
        reference's own metadata dict (vector_store.py:414-463), the scan then only consults the
        bitmap for rows that already beat the running top-k threshold."""
        key = json.dumps(filter_metadata, sort_keys=True, default=str)
        cached = self._allow_cache.get(key)
        if cached is not None and cached[0] == self._version:
            return cached[1]
        maps = []
        for s in range(self.num_shards):
            order = np.concatenate(self._row_gids[s]) if self._row_gids[s] else np.empty(0, np.uint32)
            if self.dist.world > 1:
                order = order[self.dist.rank::self.dist.world]   # this rank's stripe, in local row order
            ok = np.fromiter((self._matches_filter(self._id_of(int(g)), filter_metadata) for g in order),
                             dtype=bool, count=order.shape[0])
            pad = (-ok.shape[0]) % 32
            bits = np.packbits(np.concatenate([ok, np.zeros(pad, bool)]), bitorder="little")
            maps.append(bits.view(np.uint32) if bits.size else np.zeros(0, np.uint32))
        if len(self._allow_cache) > 16:
            self._allow_cache.clear()
        self._allow_cache[key] = (self._version, maps)
        return maps

    def _search_prefiltered(self, q: np.ndarray, limit: int, threshold: float, filter_metadata: Dict[str, Any]):
        if q.shape != (self.vector_dim,):
            raise ValueError(f"Vector dimension mismatch: expected {self.vector_dim}, got {q.shape[-1]}")
        k = min(int(limit), self.count(), _lib.MAX_K)
        if k <= 0:
            return [[]]
        floor = float(threshold) if threshold > 0 else float("-inf")
        # the bitmaps cover exactly the rows the segments hold NOW: keep writers out until the engine has consumed them
        # (the engine takes bitmap pointers without lengths; the ctypes call releases the GIL, not this lock)
        with self._lock:
            maps = self._allow_bitmaps(filter_metadata)
            if self.dist.world == 1:
                scores, gids, counts = self.engine.search_filtered_host(q[None, :], k, self.metric, floor, maps)
            else:   # SPMD: every rank passes the bitmaps of ITS rows; the ranks merge on the device (NVLink exchange)
                scores, gids, counts = self.engine.search_filtered_host(q[None, :], k, self.metric, floor, maps, exchange=True)
        if (counts < 0).any():
            raise RuntimeError("fused exchange timed out: a peer rank did not join the collective search")
        c = int(counts[0])
        return [[(self._id_of(int(g)), float(s)) for g, s in zip(gids[0, :c], scores[0, :c])]]

    async def search_async(self, query_vector: List[float], limit: int = 10, threshold: float = 0.0,
                           filter_metadata: Optional[Dict[str, Any]] = None):
        """Reference: vector_store.py:355-412.  One thread-pool hop (the ctypes call releases the GIL)
        instead of one per shard."""
        if self._batcher is not None and not filter_metadata:
            # micro-batching front-end: concurrent requests share one device pass (batcher.py)
            query_np = self._query_array(query_vector)
            if query_np.shape != (self.vector_dim,):   # same outcome as the synchronous path: logged and [] (GPU_STRICT: raises)
                return self.search(query_vector, limit, threshold, None)
            return await self._batcher.submit_async(asyncio.get_running_loop(), query_np, limit, threshold)
        return await self._run(self.search, query_vector, limit, threshold, filter_metadata)

    ALL = ALL
    max_k = _lib.MAX_K

    def _log_error(self, e):
        logger.error(f"Error in B200 batched search: {e}")

    def search_batch(self, queries, limit: int = 10, metric: Optional[str] = None) -> BatchResult:
        """Additive batch entry point: [B, dim] queries -> device-merged top-``limit`` per query."""
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        if Q.ndim != 2 or Q.shape[1] != self.vector_dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.vector_dim}, got {Q.shape[-1]}")
        k = min(int(limit), max(self.count(), 1), _lib.MAX_K)
        scores, gids, counts = self._search_arrays(Q, k, ALL, metric)
        return BatchResult(gids, scores, counts, self)

    def _matches_filter(self, vector_id: str, filter_metadata: Dict[str, Any]) -> bool:
        """Mongo-style post filter, same operator set and semantics as vector_store.py:414-463."""
        metadata = self.metadata.get(vector_id, {})
        for key, value in filter_metadata.items():
            if isinstance(value, dict) and list(value.keys())[0].startswith("$"):
                op = list(value.keys())[0]
                op_value = value[op]
                present = key in metadata
                if op == "$gt":
                    if not present or metadata[key] <= op_value:
                        return False
                elif op == "$lt":
                    if not present or metadata[key] >= op_value:
                        return False
                elif op == "$gte":
                    if not present or metadata[key] < op_value:
                        return False
                elif op == "$lte":
                    if not present or metadata[key] > op_value:
                        return False
                elif op == "$in":
                    if not present or metadata[key] not in op_value:
                        return False
                elif op == "$nin":
                    if present and metadata[key] in op_value:
                        return False
                elif op == "$exists":
                    if bool(op_value) != present:
                        return False
            elif key not in metadata or metadata[key] != value:
                return False
        return True

    # ------------------------------------------------------------------ stats
    def get_stats(self) -> Dict[str, Any]:
        """Same keys as vector_store.py:670-696 plus the engine's device counters."""
        index_stats = [
            {"shard": i, "type": self.index_type, "size": ix.size(), "stats": ix.get_stats()}
            for i, ix in enumerate(self.indices)
        ]
        est = self.engine.stats() if self.engine is not None else {}
        # observability: what the last host search streamed, as an effective rate
        # over the STORED row bytes (the bf16-shadow filter path reads half of them, so this can exceed the
        # HBM peak; the kernel's own rate is in last_kernel_ms when kernel timing is on)
        ms = est.get("last_search_ms") or 0.0
        if ms > 0 and est.get("rows_total"):
            elem = 2 if self.dtype == "bf16" else 4
            stored = est["rows_total"] * est.get("dim_padded", self.vector_dim) * elem
            est["last_search_stored_gbs"] = stored / (ms * 1e-3) / 1e9
        return {
            "vector_count": self.count(),
            "metadata_count": len(self.metadata),
            "index_type": self.index_type,
            "num_shards": self.num_shards,
            "vector_dim": self.vector_dim,
            "use_gpu": self.use_gpu,
            "indices": index_stats,
            "gpu": {
                "world_size": self.dist.world,
                "rank": self.dist.rank,
                "device": self.devices[0],
                "devices": list(self.devices),
                "metric": self.metric,
                "dtype": self.dtype,
                "engine": est,
                "shard_allocation": self.shard_map.allocation(self._shard_count),
            },
        }
