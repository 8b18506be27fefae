"""Thin Python handle over the C ABI (one engine = the row partitions of one GPU)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import ALL_SEGMENTS, EACH_SEGMENT, B200Error, DTYPES, METRICS, check


def _metric_code(metric) -> int:
    if isinstance(metric, int):
        return metric
    try:
        return METRICS[str(metric).lower()]
    except KeyError:
        raise ValueError(f"unknown metric {metric!r}; expected one of {sorted(METRICS)}") from None


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def new_out(B: int, k: int, dev):
    """Result tensors of one search as views of ONE packed device buffer
    (keys i64 [B,k] | gids i64 [B,k] | scores f32 [B,k] | counts i32 [B]) so that the host path
    needs a single device-to-host copy (``unpack_out``)."""
    import torch

    n = B * k
    packed = torch.empty(n * 20 + B * 4, dtype=torch.uint8, device=dev)
    return {
        "packed": packed,
        "keys": packed[: n * 8].view(torch.int64).view(B, k),
        "gids": packed[n * 8: n * 16].view(torch.int64).view(B, k),
        "scores": packed[n * 16: n * 20].view(torch.float32).view(B, k),
        "counts": packed[n * 20:].view(torch.int32),
    }


def unpack_out(out: Dict, B: int, k: int):
    """One D2H copy of a ``new_out`` result -> (scores, gids, counts) numpy arrays."""
    host = out["packed"].cpu().numpy()
    n = B * k
    gids = host[n * 8: n * 16].view(np.int64).reshape(B, k)
    scores = host[n * 16: n * 20].view(np.float32).reshape(B, k)
    counts = host[n * 20:].view(np.int32)
    return scores, gids, counts


class Engine:
    """Device-resident exact-search engine for ONE GPU.

    Segments are the logical WDBX shards (``VectorStore.indices[i]``, wdbx/core/vector_store.py:111-134);
    rows carry a global insertion id (gid) that the caller maps back to string ids.
    """

    def __init__(self, device: int = 0, dim: int = 384, dtype: str = "fp32", num_segments: int = 1):
        self._lib = _lib.load_library()
        self.device, self.dim, self.num_segments = int(device), int(dim), int(num_segments)
        self.dtype = str(dtype).lower()
        if self.dtype not in DTYPES:
            raise ValueError(f"unknown dtype {dtype!r}")
        h = C.c_void_p()
        check(self._lib.wdbx_b200_create(self.device, self.dim, DTYPES[self.dtype], self.num_segments, C.byref(h)))
        self._h = h

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.wdbx_b200_destroy(self._h)
            self._h = None

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise B200Error(_lib.ERR_ARG, "engine is closed")
        return self._h

    # ------------------------------------------------------------------ mutation
    def reserve(self, segment: int, rows: int):
        check(self._lib.wdbx_b200_reserve(self._handle(), segment, rows))

    def append(self, segment: int, rows, gids: Optional[np.ndarray] = None) -> int:
        """Append [n, dim] fp32 rows (numpy array, or a CUDA torch tensor on this device).
        Returns the segment-local index of the first appended row."""
        first = C.c_int64(-1)
        g = None
        if gids is not None:
            g = np.ascontiguousarray(gids, dtype=np.uint32)
        if isinstance(rows, np.ndarray) or not hasattr(rows, "data_ptr"):
            a = np.ascontiguousarray(rows, dtype=np.float32)
            if a.ndim == 1:
                a = a[None, :]
            if a.ndim != 2 or a.shape[1] != self.dim:
                raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {a.shape[-1]}")
            if g is not None and g.shape[0] != a.shape[0]:
                raise ValueError("gids length mismatch")
            check(self._lib.wdbx_b200_append(self._handle(), segment, _np_ptr(a), a.shape[0], 0, _np_ptr(g), C.byref(first)))
        else:
            import torch

            t = rows
            if t.dim() == 1:
                t = t[None, :]
            if not t.is_cuda or t.device.index != self.device:
                raise ValueError("tensor rows must live on the engine's CUDA device")
            if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != self.dim:
                raise ValueError(f"Vector dimension mismatch: expected fp32 [n, {self.dim}]")
            t = t.contiguous()
            if g is not None and g.shape[0] != t.shape[0]:
                raise ValueError("gids length mismatch")
            torch.cuda.current_stream(t.device).synchronize()  # the ingest kernel runs on the engine's own stream
            check(self._lib.wdbx_b200_append(self._handle(), segment, C.c_void_p(t.data_ptr()), t.shape[0], 1,
                                             _np_ptr(g), C.byref(first)))
        return int(first.value)

    def overwrite(self, segment: int, row: int, vector):
        v = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        if v.shape[0] != self.dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {v.shape[0]}")
        check(self._lib.wdbx_b200_overwrite(self._handle(), segment, row, _np_ptr(v)))

    def tombstone(self, segment: int, row: int, dead: bool = True):
        check(self._lib.wdbx_b200_tombstone(self._handle(), segment, row, 1 if dead else 0))

    def clear(self, segment: int = ALL_SEGMENTS):
        check(self._lib.wdbx_b200_clear(self._handle(), segment))

    def read_row(self, segment: int, row: int) -> np.ndarray:
        out = np.empty(self.dim, dtype=np.float32)
        check(self._lib.wdbx_b200_read_row(self._handle(), segment, row, _np_ptr(out)))
        return out

    def read_rows(self, segment: int, row0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float32)
        if n > 0:
            check(self._lib.wdbx_b200_read_rows(self._handle(), segment, row0, n, _np_ptr(out)))
        return out

    # ------------------------------------------------------------------ search
    def upload(self, queries):
        """Host fp32 array -> CUDA tensor on this engine's device (queries of the SPMD path)."""
        import torch

        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        return torch.from_numpy(q).to(torch.device("cuda", self.device), non_blocking=False)

    def search_host(self, queries, k: int, metric="cosine", per_segment: bool = False,
                    want_keys: bool = False, segment: int = ALL_SEGMENTS):
        """Host-buffer search (H2D + kernels + D2H inside the call).

        Returns (scores, gids, counts[, keys]); shapes [B, k] / [B], or with ``per_segment``
        [num_segments, B, k] / [num_segments, B]."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {q.shape[-1]}")
        B = q.shape[0]
        lead = (self.num_segments, B) if per_segment else (B,)
        scores = np.empty(lead + (k,), dtype=np.float32)
        gids = np.empty(lead + (k,), dtype=np.int64)
        counts = np.empty(lead, dtype=np.int32)
        keys = np.empty(lead + (k,), dtype=np.uint64) if want_keys else None
        sel = EACH_SEGMENT if per_segment else int(segment)
        check(self._lib.wdbx_b200_search_host(self._handle(), sel, _np_ptr(q), B, k,
                                              _metric_code(metric), _np_ptr(scores), _np_ptr(gids), _np_ptr(keys),
                                              _np_ptr(counts)))
        return (scores, gids, counts, keys) if want_keys else (scores, gids, counts)

    def search_exchange_host(self, queries, k: int, metric="cosine"):
        """Collective host-buffer search (every rank, same queries): pinned H2D -> local exact top-k ->
        on-device NVLink exchange + merge -> one D2H.  Returns (scores, gids, counts) of the GLOBAL top-k;
        counts == -1 means a peer did not join in time."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {q.shape[-1]}")
        B = q.shape[0]
        scores = np.empty((B, k), dtype=np.float32)
        gids = np.empty((B, k), dtype=np.int64)
        counts = np.empty((B,), dtype=np.int32)
        check(self._lib.wdbx_b200_search_exchange_host(self._handle(), _np_ptr(q), B, k, _metric_code(metric),
                                                       _np_ptr(scores), _np_ptr(gids), None, _np_ptr(counts)))
        return scores, gids, counts

    def search_filtered_host(self, queries, k: int, metric="cosine", min_score: float = float("-inf"), allow=None,
                             exchange: bool = False):
        """Opt-in pre-filtered search: ``allow`` is a list (one entry per segment) of uint32 bitmaps over the
        segment's rows (bit 1 = row may be returned; None = all rows), ``min_score`` a score floor.
        Returns (scores, gids, counts) of the exact top-k among the allowed rows, merged over segments.
        ``exchange``: COLLECTIVE form (every rank passes the bitmaps of its own rows; on-device NVLink merge)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"Vector dimension mismatch: expected {self.dim}, got {q.shape[-1]}")
        B = q.shape[0]
        scores = np.empty((B, k), dtype=np.float32)
        gids = np.empty((B, k), dtype=np.int64)
        counts = np.empty((B,), dtype=np.int32)
        ptrs = None
        keep = []
        if allow is not None:
            if len(allow) != self.num_segments:
                raise ValueError("allow must have one entry per segment")
            arr = (C.c_void_p * self.num_segments)()
            for s, bm in enumerate(allow):
                if bm is None:
                    arr[s] = None
                else:
                    b = np.ascontiguousarray(bm, dtype=np.uint32)
                    keep.append(b)
                    arr[s] = b.ctypes.data
            ptrs = C.cast(arr, C.c_void_p)
        fn = self._lib.wdbx_b200_search_exchange_filtered_host if exchange else self._lib.wdbx_b200_search_filtered_host
        check(fn(self._handle(), _np_ptr(q), B, k, _metric_code(metric), C.c_float(min_score), ptrs, _np_ptr(scores),
                 _np_ptr(gids), _np_ptr(counts)))
        return scores, gids, counts

    def search(self, q_dev, k: int, metric="cosine", segment: int = ALL_SEGMENTS, out: Optional[Dict] = None,
               stream=None) -> Dict:
        """Device-resident search: q_dev is a CUDA fp32 tensor [B, dim]; outputs are CUDA tensors
        (allocated here unless ``out`` carries preallocated ones -- required under graph capture).
        Asynchronous on the current (or given) torch stream."""
        import torch

        if q_dev.dim() == 1:
            q_dev = q_dev[None, :]
        if not q_dev.is_cuda or q_dev.dtype != torch.float32 or q_dev.shape[1] != self.dim or not q_dev.is_contiguous():
            raise ValueError(f"q_dev must be a contiguous CUDA fp32 tensor [B, {self.dim}]")
        B = q_dev.shape[0]
        dev = q_dev.device
        if out is None:
            out = new_out(B, k, dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        p = lambda name: C.c_void_p(out[name].data_ptr()) if out.get(name) is not None else None  # noqa: E731
        check(self._lib.wdbx_b200_search(self._handle(), segment, C.c_void_p(q_dev.data_ptr()), B, k,
                                         _metric_code(metric), p("keys"), p("scores"), p("gids"), p("counts"),
                                         C.c_void_p(st.cuda_stream)))
        return out

    # ------------------------------------------------------------------ fused cross-GPU exchange
    exchange_ready = False
    XCHG_MAX_B, XCHG_MAX_K = 8, 128

    def exchange_setup(self, rank: int, world: int, all_gather_bytes) -> None:
        """Create this rank's peer-mapped exchange buffer and attach every peer's.
        ``all_gather_bytes(b: bytes) -> List[bytes]`` gathers one 64-byte IPC handle per rank
        (rank order); the host layer implements it with torch.distributed."""
        h = (C.c_ubyte * 64)()
        check(self._lib.wdbx_b200_exchange_init(self._handle(), rank, world, C.cast(h, C.c_void_p)))
        handles = all_gather_bytes(bytes(h))
        if len(handles) != world or any(len(x) != 64 for x in handles):
            raise ValueError("exchange_setup: expected one 64-byte handle per rank")
        blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        check(self._lib.wdbx_b200_exchange_attach(self._handle(), world, C.cast(blob, C.c_void_p)))
        self.exchange_ready = True

    def search_exchange(self, q_dev, k: int, metric="cosine", out: Optional[Dict] = None, stream=None) -> Dict:
        """COLLECTIVE fused search: scan + NVLink key push + global merge in ONE kernel per rank."""
        import torch

        if q_dev.dim() == 1:
            q_dev = q_dev[None, :]
        if not q_dev.is_cuda or q_dev.dtype != torch.float32 or q_dev.shape[1] != self.dim or not q_dev.is_contiguous():
            raise ValueError(f"q_dev must be a contiguous CUDA fp32 tensor [B, {self.dim}]")
        B = q_dev.shape[0]
        dev = q_dev.device
        if out is None:
            out = new_out(B, k, dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        p = lambda name: C.c_void_p(out[name].data_ptr()) if out.get(name) is not None else None  # noqa: E731
        check(self._lib.wdbx_b200_search_exchange(self._handle(), C.c_void_p(q_dev.data_ptr()), B, k,
                                                  _metric_code(metric), p("keys"), p("scores"), p("gids"), p("counts"),
                                                  C.c_void_p(st.cuda_stream)))
        return out

    def merge(self, keys, out: Optional[Dict] = None, stream=None) -> Dict:
        """k-way merge of keys [G, B, k] (int64 view of the packed u64 keys) on the device."""
        import torch

        if keys.dim() != 3 or keys.dtype != torch.int64 or not keys.is_cuda or not keys.is_contiguous():
            raise ValueError("keys must be a contiguous CUDA int64 tensor [G, B, k]")
        G, B, k = keys.shape
        dev = keys.device
        if out is None:
            out = new_out(B, k, dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        p = lambda name: C.c_void_p(out[name].data_ptr()) if out.get(name) is not None else None  # noqa: E731
        check(self._lib.wdbx_b200_merge(self._handle(), C.c_void_p(keys.data_ptr()), G, B, k, p("keys"), p("scores"),
                                        p("gids"), p("counts"), C.c_void_p(st.cuda_stream)))
        return out

    # ------------------------------------------------------------------ misc
    def set_tuning(self, warps: int = 0, stages: int = 0, rows_unroll: int = 0, grid: int = 0, evict_first: int = -1):
        check(self._lib.wdbx_b200_set_tuning(self._handle(), warps, stages, rows_unroll, grid, evict_first))

    def set_option(self, name: str, value: int):
        """Routing knob of a live engine ("shadow_min_mb", "gemm_min_batch", "gemm_mode", "pdl", "overlap", "filter_i8",
        "queries_per_pass");
        results are identical for every setting.  Benchmark hook."""
        check(self._lib.wdbx_b200_set_option(self._handle(), name.encode(), int(value)))

    def set_kernel_timing(self, enable: bool = True):
        """Bracket the dominant kernel of every search with CUDA events; `stats()` then carries
        `last_kernel` (1 = K1 scan, 2 = K2b filter over the bf16 shadow, 3 = small-batch filter over the int8 shadow) and
        `last_kernel_ms`.  Measurement hook."""
        check(self._lib.wdbx_b200_set_kernel_timing(self._handle(), 1 if enable else 0))

    def stats(self) -> Dict:
        st = _lib.Stats()
        check(self._lib.wdbx_b200_get_stats(self._handle(), C.byref(st)))
        d = {name: getattr(st, name) for name, _ in _lib.Stats._fields_ if not name.startswith(("seg_", "reserved"))}
        d["seg_rows"] = list(st.seg_rows)[: self.num_segments]
        d["seg_live"] = list(st.seg_live)[: self.num_segments]
        return d


def device_count() -> int:
    """Number of visible CUDA devices; raises B200Error when CUDA is unusable."""
    return check(_lib.load_library().wdbx_b200_device_count())
