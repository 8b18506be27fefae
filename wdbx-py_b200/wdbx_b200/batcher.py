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


class _LoopFuture:
    """An asyncio future seen from the batcher thread: results are collected per event loop and handed over in one call."""
    __slots__ = ("loop", "fut", "outcome")

    def __init__(self, loop, fut):
        self.loop, self.fut, self.outcome = loop, fut, None

    def cancelled(self):
        return self.fut.cancelled()

    def done(self):
        return self.outcome is not None or self.fut.done()

    def set_result(self, result):
        self.outcome = (result, None)

    def set_exception(self, error):
        self.outcome = (None, error)


def _complete_many(items):
    for fut, (result, error) in items:
        if not fut.done():               # cancelled meanwhile (client gone)
            if error is not None:
                fut.set_exception(error)
            else:
                fut.set_result(result)


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

    def submit_async(self, loop, query: np.ndarray, limit: int, threshold: float):
        """The same for a coroutine: an asyncio future of `loop`.  All futures of one pass that belong to one loop are
        completed by ONE ``call_soon_threadsafe`` (one wake-up of the loop per pass instead of one per request)."""
        fut = loop.create_future()
        self._q.put((query, int(limit), float(threshold), _LoopFuture(loop, fut)))
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
        batch = [b for b in batch if not b[3].cancelled()]
        if not batch:
            return
        try:
            self._answer(batch)
        finally:
            # hand the asyncio futures of this pass to their loops: one thread-safe call per loop
            per_loop = {}
            for _, _, _, fut in batch:
                if isinstance(fut, _LoopFuture) and fut.outcome is not None:
                    per_loop.setdefault(fut.loop, []).append((fut.fut, fut.outcome))
            for loop, items in per_loop.items():
                try:
                    loop.call_soon_threadsafe(_complete_many, items)
                except RuntimeError:     # the loop is closed: nobody is waiting any more
                    pass

    def _answer(self, batch):
        store = self._store
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
            # one conversion per array instead of one numpy scalar per element (this loop is the per-request cost)
            gl, sl, cl = gids.tolist(), scores.tolist(), counts.tolist()
            id_of, meta = store._id_of, store.metadata
            for i, (_, limit, threshold, fut) in enumerate(batch):
                c = max(0, min(cl[i], limit))     # limit <= 0 -> [] like the synchronous path (never a negative slice)
                res = [(id_of(g), s) for g, s in zip(gl[i][:c], sl[i][:c])]
                if threshold > 0:
                    res = [r for r in res if r[1] >= threshold]
                self._deliver(fut, [(vid, sc, meta.get(vid, {})) for vid, sc in res])
        except Exception as e:  # reference convention: log + [] unless strict (indexing.py:1028-1030)
            if not store.strict:
                store._log_error(e)
            for _, _, _, fut in batch:
                if not fut.done():
                    self._deliver(fut, [], e if store.strict else None)
