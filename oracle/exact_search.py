"""CPU oracle for WDBX's exact vector_search path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path (``wdbx-py_b200/``)
never imports anything under ``oracle/`` and has no CPU fallback.

What is restated (reference = /root/reference, donaldfilimon/wdbx-py):

* ``normalize``            <- ``FaissIndex._normalize_vector``  wdbx/core/indexing.py:851-856
* ``flat_ip_search``       <- ``faiss.IndexFlatIP.search`` as called at wdbx/core/indexing.py:1013.
  The arithmetic lives in the un-vendored, unpinned third-party wheel ``faiss-cpu>=1.7.0``
  (requirements.txt:20).  Its published algorithm for IndexFlatIP: for every stored row compute the
  fp32 inner product with the query, keep the k largest in a heap, return them best-first;
  fewer than k rows => padded with label -1.
* ``faiss_index_search``   <- ``FaissIndex.search``              wdbx/core/indexing.py:983-1030
* ``store_search``         <- ``VectorStore.search``             wdbx/core/vector_store.py:301-353
* ``matches_filter``       <- ``VectorStore._matches_filter``    wdbx/core/vector_store.py:414-463

Pinning: this restatement is checked against (a) the golden vectors produced by running the
reference's own unmodified ``FaissIndex`` / ``HNSWIndex`` / ``VectorStore`` code on top of
exact stand-ins for the two absent third-party libraries (tests/golden/make_golden.py,
fixtures committed under tests/golden/), and (b) the reference's own test fixtures
(tests/test_core.py:135-142, :199-230).  The third-party arithmetic itself (summation order
inside faiss) is not reproducible here, so fp32 scores are compared within the tolerance of
SURVEY.md section 8c and ranks are adjudicated in fp64.

Extensions (the reference implements cosine only): ``ip`` = raw inner product,
``l2`` = negative squared euclidean distance (larger is better, so the facade's
"sort descending" contract of vector_store.py:330 still holds).

Tie rule (ours; the reference's is library-defined): higher score first, equal score =>
lower global insertion row first.  NaN scores rank as -inf.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

METRICS = ("cosine", "ip", "l2")


def normalize(v: np.ndarray) -> np.ndarray:
    """v / ||v||_2 in fp32, zero vector passes through (indexing.py:851-856)."""
    v = np.asarray(v, dtype=np.float32)
    norm = np.linalg.norm(v)
    if norm > 0:
        return v / norm
    return v


def normalize_rows(X: np.ndarray) -> np.ndarray:
    """Row-wise ``normalize`` (batch_add normalises row by row, indexing.py:937-942)."""
    X = np.asarray(X, dtype=np.float32)
    n = np.linalg.norm(X, axis=1).astype(np.float32)
    out = X.copy()
    nz = n > 0
    out[nz] = X[nz] / n[nz, None]
    return out


def bf16_round(X: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) -> fp32.  bf16 storage is part of the data
    (SURVEY.md section 8c): the oracle scores the rounded matrix in fp32/fp64."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    u = X.view(np.uint32).astype(np.uint64)
    nan = np.isnan(X)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32).reshape(X.shape).copy()
    out[nan] = np.nan
    return out


def scores_fp32(X: np.ndarray, q: np.ndarray, metric: str = "cosine") -> np.ndarray:
    """fp32 score of every row of X against q.  cosine follows the reference exactly:
    normalised rows (indexing.py:886) dotted with the normalised query (indexing.py:1002)."""
    X = np.asarray(X, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    if metric == "cosine":
        return normalize_rows(X) @ normalize(q)
    if metric == "ip":
        return X @ q
    if metric == "l2":
        d = X - q[None, :]
        return -np.einsum("ij,ij->i", d, d, dtype=np.float32)
    raise ValueError(f"unknown metric {metric!r}")


def scores_fp64(X: np.ndarray, q: np.ndarray, metric: str = "cosine") -> np.ndarray:
    """fp64 adjudicator for rank comparisons."""
    X = np.asarray(X, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    if metric == "cosine":
        xn = np.linalg.norm(X, axis=1)
        qn = np.linalg.norm(q)
        dots = X @ q
        with np.errstate(divide="ignore", invalid="ignore"):
            s = dots / (xn * qn)
        s[(xn == 0) | (qn == 0)] = 0.0
        return s
    if metric == "ip":
        return X @ q
    if metric == "l2":
        d = X - q[None, :]
        return -np.einsum("ij,ij->i", d, d)
    raise ValueError(f"unknown metric {metric!r}")


def topk_desc(scores: np.ndarray, k: int, rows: Optional[np.ndarray] = None,
              dead: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """k best by (score desc, row asc); NaN -> -inf; ``dead`` rows are excluded.
    Returns (row_ids int64 [k'], scores [k']) with k' = min(k, #live)."""
    s = np.array(scores, copy=True)
    s[np.isnan(s)] = -np.inf
    n = s.shape[0]
    r = np.arange(n, dtype=np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
    if dead is not None:
        live = ~np.asarray(dead, dtype=bool)
        s, r = s[live], r[live]
    k = min(int(k), s.shape[0])
    if k <= 0:
        return np.empty(0, np.int64), np.empty(0, s.dtype)
    if k < s.shape[0]:
        # keep everything tied with the k-th value so the (row asc) tie-break is exact
        kth = np.partition(s, s.shape[0] - k)[s.shape[0] - k]
        cand = np.flatnonzero(s >= kth)
    else:
        cand = np.arange(s.shape[0])
    order = np.lexsort((r[cand], -s[cand]))[:k]
    sel = cand[order]
    return r[sel], s[sel]


def flat_ip_search(Xhat: np.ndarray, qhat: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """``IndexFlatIP.search`` for one query: (D[k], I[k]) best-first, padded with -1 labels
    when fewer than k rows exist (the padding the reference filters at indexing.py:1023)."""
    Xhat = np.asarray(Xhat, dtype=np.float32).reshape(-1, qhat.shape[-1])
    s = Xhat @ np.asarray(qhat, dtype=np.float32)
    idx, val = topk_desc(s, k)
    D = np.full(k, -np.finfo(np.float32).max, dtype=np.float32)
    I = np.full(k, -1, dtype=np.int64)
    D[: len(idx)] = val
    I[: len(idx)] = idx
    return D, I


def faiss_index_search(Xhat: np.ndarray, index_to_id: Dict[int, str], q: np.ndarray,
                       limit: int) -> List[Tuple[str, float]]:
    """``FaissIndex.search`` (indexing.py:983-1030) over an already-normalised matrix."""
    n = Xhat.shape[0]
    if n == 0:
        return []
    qhat = normalize(np.asarray(q, dtype=np.float32))
    actual = min(limit, n)
    if actual == 0:
        return []
    D, I = flat_ip_search(Xhat, qhat, actual)
    return [(index_to_id.get(int(i), str(i)), float(s)) for i, s in zip(I, D) if i != -1]


def matches_filter(metadata: Dict[str, Any], flt: Dict[str, Any]) -> bool:
    """Mongo-style post filter (vector_store.py:414-463)."""
    for key, value in flt.items():
        if isinstance(value, dict) and list(value.keys())[0].startswith("$"):
            op = list(value.keys())[0]
            ov = value[op]
            if op == "$gt":
                if key not in metadata or metadata[key] <= ov:
                    return False
            elif op == "$lt":
                if key not in metadata or metadata[key] >= ov:
                    return False
            elif op == "$gte":
                if key not in metadata or metadata[key] < ov:
                    return False
            elif op == "$lte":
                if key not in metadata or metadata[key] > ov:
                    return False
            elif op == "$in":
                if key not in metadata or metadata[key] not in ov:
                    return False
            elif op == "$nin":
                if key in metadata and metadata[key] in ov:
                    return False
            elif op == "$exists":
                if ov and key not in metadata:
                    return False
                if not ov and key in metadata:
                    return False
        else:
            if key not in metadata or metadata[key] != value:
                return False
    return True


def store_search(shard_results: Sequence[List[Tuple[str, float]]], limit: int,
                 threshold: float = 0.0,
                 filter_metadata: Optional[Dict[str, Any]] = None,
                 metadata: Optional[Dict[str, Dict[str, Any]]] = None,
                 ) -> List[Tuple[str, float, Dict[str, Any]]]:
    """``VectorStore.search`` merge (vector_store.py:323-351): concat per-shard lists in shard
    order, stable sort by score desc, threshold (only when > 0), post-filter, ``[:limit]``,
    attach metadata."""
    metadata = metadata or {}
    allr: List[Tuple[str, float]] = []
    for res in shard_results:
        allr.extend(res)
    allr.sort(key=lambda x: x[1], reverse=True)
    if threshold > 0:
        allr = [r for r in allr if r[1] >= threshold]
    if filter_metadata:
        allr = [r for r in allr if matches_filter(metadata.get(r[0], {}), filter_metadata)]
    allr = allr[:limit]
    return [(vid, sc, metadata.get(vid, {})) for vid, sc in allr]


class OracleStore:
    """Minimal end-to-end restatement: shard placement is *given* (the reference's
    ``abs(hash(id)) % num_shards`` is process-salted, vector_store.py:188-190), every shard is
    a FAISS-Flat index restated in numpy, search is ``store_search`` over the shard lists."""

    def __init__(self, dim: int, num_shards: int = 1, metric: str = "cosine"):
        self.dim, self.num_shards, self.metric = dim, num_shards, metric
        self.rows: List[List[np.ndarray]] = [[] for _ in range(num_shards)]
        self.ids: List[List[str]] = [[] for _ in range(num_shards)]
        self.gids: List[List[int]] = [[] for _ in range(num_shards)]
        self.metadata: Dict[str, Dict[str, Any]] = {}
        self._next_gid = 0

    def add(self, shard: int, vid: str, vec, meta: Optional[dict] = None):
        self.rows[shard].append(np.asarray(vec, dtype=np.float32))
        self.ids[shard].append(vid)
        self.gids[shard].append(self._next_gid)
        self._next_gid += 1
        self.metadata[vid] = meta or {}

    def shard_search(self, shard: int, q, limit: int) -> List[Tuple[str, float]]:
        if not self.rows[shard]:
            return []
        X = np.stack(self.rows[shard])
        s = scores_fp32(X, np.asarray(q, np.float32), self.metric)
        idx, val = topk_desc(s, limit, rows=np.asarray(self.gids[shard]))
        g2i = dict(zip(self.gids[shard], self.ids[shard]))
        return [(g2i[int(g)], float(v)) for g, v in zip(idx, val)]

    def search(self, q, limit=10, threshold=0.0, filter_metadata=None):
        per_shard = [self.shard_search(s, q, limit) for s in range(self.num_shards)]
        return store_search(per_shard, limit, threshold, filter_metadata, self.metadata)


# ---------------------------------------------------------------------------------------------
# Comparison helpers used by the parity tests (tolerances of SURVEY.md section 8c).

def score_tolerance(s64: np.ndarray, qnorm: float, xnorm: np.ndarray, metric: str) -> np.ndarray:
    """abs tolerance: 1e-5*|s| + 1e-6*||q||*||x|| (cosine: the operands are unit vectors)."""
    if metric == "cosine":
        floor = 1e-6
    elif metric == "ip":
        floor = 1e-6 * qnorm * xnorm
    else:  # l2: cancellation scale is ||q||^2 + ||x||^2
        floor = 1e-6 * (qnorm * qnorm + xnorm * xnorm)
    return 1e-5 * np.abs(s64) + floor


def rank_tie_window(dim: int, qnorm: float, xnorm: float, metric: str) -> float:
    """fp64 score gap under which an id mismatch counts as a tie: 8*sqrt(D)*2^-24*||q||*||x||."""
    scale = 1.0 if metric == "cosine" else (qnorm * xnorm if metric == "ip"
                                            else qnorm * qnorm + xnorm * xnorm)
    return 8.0 * np.sqrt(dim) * 2.0 ** -24 * scale


def check_topk(X: np.ndarray, q: np.ndarray, metric: str, k: int, got_rows: np.ndarray,
               got_scores: np.ndarray, dead: Optional[np.ndarray] = None) -> Dict[str, Any]:
    """Compare a returned top-k (global rows + fp32 scores) with the fp64 truth.
    Returns recall, #entries that differ only inside the tie window, max score error / tol."""
    s64 = scores_fp64(X, q, metric)
    s64[np.isnan(s64)] = -np.inf
    want_rows, want_s = topk_desc(s64, k, dead=dead)
    got_rows = np.asarray(got_rows, dtype=np.int64)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    out: Dict[str, Any] = {"count_ok": len(got_rows) == len(want_rows)}
    qn = float(np.linalg.norm(np.asarray(q, np.float64)))
    xn_all = np.linalg.norm(np.asarray(X, np.float64), axis=1)
    in_window = 0
    hard_miss = 0
    for j in range(min(len(got_rows), len(want_rows))):
        if got_rows[j] == want_rows[j]:
            continue
        w = rank_tie_window(X.shape[1], qn, float(max(xn_all[got_rows[j]], xn_all[want_rows[j]])), metric)
        if abs(s64[got_rows[j]] - want_s[j]) <= w:
            in_window += 1
        else:
            hard_miss += 1
    out["in_tie_window"] = in_window
    out["hard_mismatch"] = hard_miss
    out["recall"] = (len(set(got_rows.tolist()) & set(want_rows.tolist())) / max(1, len(want_rows)))
    finite = np.isfinite(s64[got_rows]) if len(got_rows) else np.zeros(0, bool)
    if finite.any():
        gr = got_rows[finite]
        err = np.abs(got_scores[finite] - s64[gr])
        tol = score_tolerance(s64[gr], qn, xn_all[gr], metric)
        out["max_err_over_tol"] = float(np.max(err / tol))
        out["max_abs_err"] = float(np.max(err))
    else:
        out["max_err_over_tol"] = 0.0
        out["max_abs_err"] = 0.0
    out["sorted"] = bool(np.all(np.diff(got_scores) <= 0)) if len(got_scores) > 1 else True
    return out
