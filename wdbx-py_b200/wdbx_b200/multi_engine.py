"""Several GPUs behind ONE ordinary process (north_star: "the shard manager maps num_shards onto the 8 GPUs of
one box ... the plugin, API-server and CLI layers are untouched").

``MultiEngine`` has the interface ``VectorStore`` expects from ``Engine`` but owns G engines, one per device.
Rows of every segment (= WDBX shard) are striped over them -- position n of a segment lives on device
``n % G`` at local row ``n // G``, the single-process form of ``shard_map.ShardMap`` and the replacement of
``ShardManager._allocate_shards`` (wdbx/core/distributed.py:547-654) -- so any ``num_shards`` balances on any G.
Ingest / CRUD are per-engine calls routed here; a search is ONE C call (``wdbx_b200_group_search_host``) that
launches all G devices and lets them merge their top-k over NVLink on the device (include/wdbx_b200.h).
``WDBX(enable_gpu=True, config={"GPU_DEVICES": "0-7"})`` therefore serves ``vector_search`` /
``vector_search_async`` (micro-batcher on) from the REST server (wdbx/api/server.py:141-152) or the CLI
(wdbx/cli.py:541) without ``torchrun``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ALL_SEGMENTS, EACH_SEGMENT, check
from .engine import Engine, _metric_code, _np_ptr, new_out


def parse_devices(spec) -> Optional[List[int]]:
    """``GPU_DEVICES``: None, "all", "0-7", "0,2,3", 4 (= the first four) or a list of ints."""
    if spec is None or spec == "" or spec is False:
        return None
    if isinstance(spec, (list, tuple)):
        return [int(d) for d in spec]
    if isinstance(spec, int):
        return list(range(spec))
    s = str(spec).strip().lower()
    if s == "all":
        from .engine import device_count

        return list(range(device_count()))
    out: List[int] = []
    for part in s.split(","):
        part = part.strip()
        if "-" in part:
            a, b = part.split("-", 1)
            out.extend(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    if len(set(out)) != len(out):
        raise ValueError(f"GPU_DEVICES lists a device twice: {spec!r}")
    return out


class _CGroup:
    """ctypes handle of a wdbx_b200_group (the engines stay owned by the MultiEngine)."""

    def __init__(self, engines: Sequence[Engine]):
        self._lib = _lib.load_library()
        arr = (C.c_void_p * len(engines))(*[e._handle() for e in engines])
        h = C.c_void_p()
        check(self._lib.wdbx_b200_group_create(arr, len(engines), C.byref(h)))
        self._h = h
        self.n, self.num_segments, self.dim = len(engines), engines[0].num_segments, engines[0].dim

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wdbx_b200_group_destroy(self._h)
            self._h = None

    def search_host(self, q: np.ndarray, k: int, metric: int, segment: int, min_score: float, allow, want_keys: bool):
        B = q.shape[0]
        lead = (self.num_segments, B) if segment == EACH_SEGMENT else (B,)
        scores = np.empty(lead + (k,), dtype=np.float32)
        gids = np.empty(lead + (k,), dtype=np.int64)
        counts = np.empty(lead, dtype=np.int32)
        keys = np.empty(lead + (k,), dtype=np.uint64) if want_keys else None
        ptrs, keep = None, []
        if allow is not None:   # engine-major [n * num_segments] bitmaps over each engine's own rows
            arr = (C.c_void_p * (self.n * self.num_segments))()
            for i, per_engine in enumerate(allow):
                for s, bm in enumerate(per_engine):
                    if bm is not None:
                        b = np.ascontiguousarray(bm, dtype=np.uint32)
                        keep.append(b)
                        arr[i * self.num_segments + s] = b.ctypes.data
            ptrs = C.cast(arr, C.c_void_p)
        check(self._lib.wdbx_b200_group_search_host(self._h, int(segment), _np_ptr(q), B, k, metric, C.c_float(min_score),
                                                    ptrs, _np_ptr(scores), _np_ptr(gids), _np_ptr(keys), _np_ptr(counts)))
        return (scores, gids, counts, keys) if want_keys else (scores, gids, counts)

    def search_device(self, q_dev, k: int, metric: int, out: Dict, stream) -> Dict:
        p = lambda name: C.c_void_p(out[name].data_ptr()) if out.get(name) is not None else None  # noqa: E731
        check(self._lib.wdbx_b200_group_search(self._h, C.c_void_p(q_dev.data_ptr()), q_dev.shape[0], k, metric, p("keys"),
                                               p("scores"), p("gids"), p("counts"), C.c_void_p(stream.cuda_stream)))
        return out


class MultiEngine:
    XCHG_MAX_B, XCHG_MAX_K = Engine.XCHG_MAX_B, Engine.XCHG_MAX_K

    def __init__(self, devices: Sequence[int], dim: int, dtype: str = "fp32", num_segments: int = 1, *,
                 _engine_factory=None, _group_factory=None):   # test seams (tests/ inject numpy doubles)
        devices = [int(d) for d in devices]
        if len(devices) < 2:
            raise ValueError("MultiEngine needs at least two devices")
        if len(set(devices)) != len(devices):
            raise ValueError(f"devices must be distinct: {devices}")
        self.devices, self.dim, self.num_segments = devices, int(dim), int(num_segments)
        self.dtype = str(dtype).lower()
        self.device = devices[0]
        make = _engine_factory if _engine_factory is not None else Engine
        self.engines = []
        try:
            for d in devices:
                self.engines.append(make(d, self.dim, self.dtype, self.num_segments))
            self._group = (_group_factory or _CGroup)(self.engines)
        except Exception:
            self.close()
            raise
        self._seg_rows = [0] * self.num_segments   # positions handed out per segment
        self._next_gid = 0

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        g = getattr(self, "_group", None)
        if g is not None:
            g.close()
            self._group = None
        for e in getattr(self, "engines", []):
            e.close()
        self.engines = []

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    @property
    def G(self) -> int:
        return len(self.devices)

    def _owner(self, row: int):
        return self.engines[row % self.G], row // self.G

    # ------------------------------------------------------------------ mutation
    def reserve(self, segment: int, rows: int):
        per = (int(rows) + self.G - 1) // self.G
        for e in self.engines:
            e.reserve(segment, per)

    def append(self, segment: int, rows, gids: Optional[np.ndarray] = None) -> int:
        """Append [n, dim] rows (numpy, or a CUDA tensor on any device); returns the segment position of the first."""
        is_tensor = hasattr(rows, "data_ptr") and not isinstance(rows, np.ndarray)
        if not is_tensor:
            rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim == 1:
            rows = rows[None, :]
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {rows.shape[-1]}")
        m = int(rows.shape[0])
        n0 = self._seg_rows[segment]
        if gids is None:
            gids = np.arange(self._next_gid, self._next_gid + m, dtype=np.uint32)
            self._next_gid += m
        gids = np.ascontiguousarray(gids, dtype=np.uint32)
        if gids.shape[0] != m:
            raise ValueError("gids length mismatch")
        G = self.G
        # all-or-nothing: make room on EVERY device first (reserve is where a device runs out of memory), so a failure
        # cannot leave some stripes appended and others not -- that would shift every later position of the segment
        for d, e in enumerate(self.engines):
            off = (d - n0) % G           # first appended row that lands on device d
            if off < m:
                e.reserve(segment, (n0 + off) // G + (m - off + G - 1) // G)
        for d, e in enumerate(self.engines):
            off = (d - n0) % G
            if off >= m:
                continue
            part = rows[off::G]
            if is_tensor:
                import torch

                part = part.to(torch.device("cuda", self.devices[d])).contiguous()
            first = e.append(segment, part, gids=gids[off::G])
            if first != (n0 + off) // G:
                raise RuntimeError(f"segment {segment}: device {self.devices[d]} row {first} != expected "
                                   f"{(n0 + off) // G} (engines out of sync)")
        self._seg_rows[segment] = n0 + m
        return n0

    def overwrite(self, segment: int, row: int, vector):
        e, local = self._owner(int(row))
        e.overwrite(segment, local, vector)

    def tombstone(self, segment: int, row: int, dead: bool = True):
        e, local = self._owner(int(row))
        e.tombstone(segment, local, dead)

    def clear(self, segment: int = ALL_SEGMENTS):
        for e in self.engines:
            e.clear(segment)
        for s in (range(self.num_segments) if segment < 0 else [segment]):
            self._seg_rows[s] = 0

    def read_row(self, segment: int, row: int) -> np.ndarray:
        e, local = self._owner(int(row))
        return e.read_row(segment, local)

    def read_rows(self, segment: int, row0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float32)
        G = self.G
        for d, e in enumerate(self.engines):
            off = (d - row0) % G
            if off >= n:
                continue
            cnt = (n - off + G - 1) // G
            out[off::G] = e.read_rows(segment, (row0 + off) // G, cnt)
        return out

    # ------------------------------------------------------------------ search
    def _queries(self, queries) -> np.ndarray:
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {q.shape[-1]}")
        return q

    def search_host(self, queries, k: int, metric="cosine", per_segment: bool = False, want_keys: bool = False,
                    segment: int = ALL_SEGMENTS):
        """Same contract as ``Engine.search_host``; every device scans its stripe, the merge happens on the devices."""
        sel = EACH_SEGMENT if per_segment else int(segment)
        return self._group.search_host(self._queries(queries), int(k), _metric_code(metric), sel, float("-inf"), None,
                                       want_keys)

    def search_filtered_host(self, queries, k: int, metric="cosine", min_score: float = float("-inf"), allow=None):
        """``allow``: one uint32 bitmap per segment over the segment's POSITIONS (None = all rows); split here into
        one bitmap per (device, segment) over the device's own rows."""
        per_engine = None
        if allow is not None:
            if len(allow) != self.num_segments:
                raise ValueError("allow must have one entry per segment")
            per_engine = [[None] * self.num_segments for _ in self.engines]
            for s, bm in enumerate(allow):
                if bm is None:
                    continue
                n = self._seg_rows[s]
                bits = np.unpackbits(np.ascontiguousarray(bm, dtype=np.uint32).view(np.uint8), bitorder="little")[:n]
                if bits.shape[0] < n:
                    raise ValueError(f"allow bitmap of segment {s} is shorter than its {n} rows")
                for d in range(self.G):
                    local = bits[d::self.G]
                    pad = (-local.shape[0]) % 32
                    packed = np.packbits(np.concatenate([local, np.zeros(pad, np.uint8)]), bitorder="little")
                    per_engine[d][s] = packed.view(np.uint32) if packed.size else np.zeros(1, np.uint32)
        return self._group.search_host(self._queries(queries), int(k), _metric_code(metric), ALL_SEGMENTS,
                                       float(min_score), per_engine, False)

    def upload(self, queries):
        import torch

        return torch.from_numpy(self._queries(queries)).to(torch.device("cuda", self.device))

    def search(self, q_dev, k: int, metric="cosine", segment: int = ALL_SEGMENTS, out: Optional[Dict] = None,
               stream=None) -> Dict:
        """Device-resident search over all segments: ``q_dev`` and the outputs live on the FIRST device."""
        import torch

        if segment != ALL_SEGMENTS:
            raise ValueError("MultiEngine.search serves all segments; use search_host for one segment")
        if q_dev.dim() == 1:
            q_dev = q_dev[None, :]
        if (not q_dev.is_cuda or q_dev.device.index != self.device or q_dev.dtype != torch.float32
                or q_dev.shape[1] != self.dim or not q_dev.is_contiguous()):
            raise ValueError(f"q_dev must be a contiguous fp32 tensor [B, {self.dim}] on cuda:{self.device}")
        if out is None:
            out = new_out(q_dev.shape[0], k, q_dev.device)
        st = stream if stream is not None else torch.cuda.current_stream(q_dev.device)
        return self._group.search_device(q_dev, int(k), _metric_code(metric), out, st)

    def merge(self, keys, out: Optional[Dict] = None, stream=None) -> Dict:
        return self.engines[0].merge(keys, out=out, stream=stream)

    # ------------------------------------------------------------------ misc
    def set_tuning(self, *args, **kw):
        for e in self.engines:
            e.set_tuning(*args, **kw)

    def set_kernel_timing(self, enable: bool = True):
        for e in self.engines:
            e.set_kernel_timing(enable)

    def set_option(self, name: str, value: int):
        for e in self.engines:
            e.set_option(name, value)

    def stats(self) -> Dict:
        per = [e.stats() for e in self.engines]
        d = dict(per[0])
        for key in ("rows_total", "rows_live", "capacity_rows", "bytes_resident", "kernel_launches"):
            if key in d:
                d[key] = sum(p.get(key, 0) for p in per)
        for key in ("seg_rows", "seg_live"):
            if key in d:
                d[key] = [sum(p[key][s] for p in per) for s in range(self.num_segments)]
        d["devices"] = list(self.devices)
        d["rows_per_device"] = [p.get("rows_total", 0) for p in per]
        return d
