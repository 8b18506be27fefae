"""Micro-batching front-end for ``vector_search_async``.

The reference serves concurrent async searches with a 4-thread pool per index
(wdbx/core/indexing.py:692, :1045-1048): every request is its own full scan.  On the GPU a pass over
the rows serves many queries at almost the cost of one (the filter kernels score 16 or 128 queries per
streamed tile: 64 queries cost 2.46 ms on 10M x 768 where one costs 2.08 ms; the streaming scan K1 scores
up to 8 per row), so concurrent single-query requests (REST ``POST /api/v1/vectors/search``
wdbx/api/server.py:141-152, CLI wdbx/cli.py:541) are coalesced for a short window (200 us, up to
``GPU_BATCH_MAX`` = 64) into ONE pass.  Callers are untouched: they still await one result list per
request.  Works with one GPU and with a single-process device group (``GPU_DEVICES``).
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import List, Tuple

import numpy as np


class MicroBatcher:
    def __init__(self, store, max_batch: int = 8, window_us: float = 200.0):
        self._store = store
        self.max_batch = max(1, int(max_batch))
        self.window_s = max(0.0, float(window_us)) * 1e-6
        self._q: "queue.Queue" = queue.Queue()
        self.batches = 0          # launches issued
        self.requests = 0         # requests served
        self._thread = threading.Thread(target=self._run, name="wdbx-b200-batcher", daemon=True)
        self._thread.start()

    def submit(self, query: np.ndarray, limit: int, threshold: float) -> Future:
        fut: Future = Future()
        self._q.put((query, int(limit), float(threshold), fut))
        return fut

    def close(self):
        self._q.put(None)
        self._thread.join(timeout=5)

    # ------------------------------------------------------------------ worker
    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            batch = [item]
            deadline = time.monotonic() + self.window_s
            while len(batch) < self.max_batch:
                left = deadline - time.monotonic()
                try:
                    nxt = self._q.get(timeout=left) if left > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._q.put(None)
                    break
                batch.append(nxt)
            self._serve(batch)

    @staticmethod
    def _deliver(fut: Future, result=None, error=None):
        """a request whose caller has gone away (cancelled future: client timeout) must not take the rest of its batch down"""
        try:
            if error is not None:
                fut.set_exception(error)
            else:
                fut.set_result(result)
        except Exception:   # InvalidStateError: cancelled / already answered
            pass

    def _serve(self, batch: List[Tuple[np.ndarray, int, float, Future]]):
        store = self._store
        batch = [b for b in batch if not b[3].cancelled()]
        if not batch:
            return
        try:
            Q = np.stack([b[0] for b in batch])
            live = store.count()
            k = min(max(b[1] for b in batch), live, store.max_k)
            if k <= 0:
                for _, _, _, fut in batch:
                    self._deliver(fut, [])
                return
            scores, gids, counts = store._search_arrays(Q, k, store.ALL)
            self.batches += 1
            self.requests += len(batch)
            for i, (_, limit, threshold, fut) in enumerate(batch):
                c = max(0, min(int(counts[i]), limit))     # limit <= 0 -> [] like the synchronous path (never a negative slice)
                res = [(store._id_of(int(g)), float(s)) for g, s in zip(gids[i, :c], scores[i, :c])]
                if threshold > 0:
                    res = [r for r in res if r[1] >= threshold]
                self._deliver(fut, [(vid, sc, store.metadata.get(vid, {})) for vid, sc in res])
        except Exception as e:  # reference convention: log + [] unless strict (indexing.py:1028-1030)
            if not store.strict:
                store._log_error(e)
            for _, _, _, fut in batch:
                if not fut.done():
                    self._deliver(fut, [], e if store.strict else None)
