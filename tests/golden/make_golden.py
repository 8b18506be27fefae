#!/usr/bin/env python
"""Generate tests/golden/reference_golden.json by running the REFERENCE's own, unmodified
``VectorStore`` / ``FaissIndex`` / ``HNSWIndex`` code (imported from /root/reference) on top of
exact stand-ins for the two third-party libraries that are absent from the image
(``faiss``, ``hnswlib``; requirements.txt:18-20, unpinned).  Run in the build container only:

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

The stand-ins implement just the published contract the reference relies on:
* ``faiss.IndexFlatIP(d).add(x) / .search(q, k) -> (D, I)``: fp32 inner products, k largest,
  best-first, label -1 padding (call sites wdbx/core/indexing.py:717, :890, :950, :1013);
* ``hnswlib.Index(space="cosine")``: ``knn_query -> (labels, distances)`` with
  ``distance = 1 - cosine`` (call sites indexing.py:271-281, :378, :445, :490), answered exactly.
Everything above them -- normalisation, limit clipping, id mapping, cross-shard concat, stable
sort, threshold, metadata post-filter, ``[:limit]``, metadata attach -- is the reference's code.

Inputs are either tiny (stored verbatim) or regenerated from a seed by the tests; the shard
placement the reference chose (``abs(hash(id)) % num_shards``, process-salted, hence
PYTHONHASHSEED=0 here) is recorded so the tests can replay it.
"""
import json
import logging
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REF = "/root/reference"
OUT = Path(__file__).resolve().parent / "reference_golden.json"


# --------------------------------------------------------------------------- stand-ins
def _topk(scores, k):
    n = scores.shape[0]
    k_eff = min(k, n)
    order = np.lexsort((np.arange(n), -scores))[:k_eff]
    return order, scores[order]


class _IndexFlatIP:
    def __init__(self, d):
        self.d = d
        self.x = np.empty((0, d), dtype=np.float32)
        self.is_trained = True

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.d)
        self.x = np.concatenate([self.x, x], axis=0)

    def search(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.d)
        D = np.full((q.shape[0], k), -3.4028235e38, dtype=np.float32)
        I = np.full((q.shape[0], k), -1, dtype=np.int64)
        for b in range(q.shape[0]):
            s = self.x @ q[b]
            idx, val = _topk(s, k)
            D[b, : len(idx)] = val
            I[b, : len(idx)] = idx
        return D, I


def _install_standins():
    faiss = types.ModuleType("faiss")
    faiss.IndexFlatIP = _IndexFlatIP
    faiss.METRIC_INNER_PRODUCT = 0
    faiss.write_index = lambda index, path: None
    faiss.read_index = lambda path: (_ for _ in ()).throw(RuntimeError("no persisted index"))
    sys.modules["faiss"] = faiss

    class _HnswIndex:
        def __init__(self, space, dim):
            assert space == "cosine"
            self.dim = dim
            self.x = {}

        def init_index(self, max_elements, ef_construction, M):
            pass

        def set_ef(self, ef):
            pass

        def add_items(self, data, ids):
            data = np.asarray(data, dtype=np.float32).reshape(-1, self.dim)
            ids = np.atleast_1d(np.asarray(ids)).tolist()
            for v, i in zip(data, ids):
                n = np.linalg.norm(v)
                self.x[int(i)] = v / n if n > 0 else v  # hnswlib cosine space normalises on insert

        def knn_query(self, q, k):
            q = np.asarray(q, dtype=np.float32).reshape(-1, self.dim)
            labels = sorted(self.x)
            X = np.stack([self.x[i] for i in labels])
            L = np.empty((q.shape[0], k), dtype=np.int64)
            Dm = np.empty((q.shape[0], k), dtype=np.float32)
            for b in range(q.shape[0]):
                n = np.linalg.norm(q[b])
                qq = q[b] / n if n > 0 else q[b]
                s = X @ qq
                idx, val = _topk(s, k)
                L[b] = np.asarray(labels)[idx]
                Dm[b] = np.float32(1.0) - val
            return L, Dm

        def save_index(self, path):
            pass

        def load_index(self, path, max_elements=0):
            raise RuntimeError("no persisted index")

    hnswlib = types.ModuleType("hnswlib")
    hnswlib.Index = _HnswIndex
    sys.modules["hnswlib"] = hnswlib


# --------------------------------------------------------------------------- cases
def _store(index_type, dim, shards, tmp):
    from wdbx.core.vector_store import VectorStore  # the reference's class, unmodified
    return VectorStore(vector_dim=dim, data_dir=Path(tmp), num_shards=shards,
                       index_type=index_type)


def _res(results):
    return [[vid, float(score), meta] for vid, score, meta in results]


def ramp_case(index_type):
    # tests/test_core.py:199-230
    vectors = {f"vec_{i}": [i / 10, (i + 1) / 10, (i + 2) / 10, (i + 3) / 10] for i in range(10)}
    metadata = {vid: {"index": i, "source": "batch_test"} for i, vid in enumerate(vectors)}
    q = [0.5, 0.6, 0.7, 0.8]
    with tempfile.TemporaryDirectory() as tmp:
        vs = _store(index_type, 4, 2, tmp)
        assert vs.batch_store(vectors, metadata) == 10
        case = {
            "index_type": index_type, "dim": 4, "num_shards": 2, "query": q,
            "vectors": vectors, "metadata": metadata,
            "placement": {vid: vs._get_shard_for_id(vid) for vid in vectors},
            "limit1": _res(vs.search(q, limit=1)),
            "limit10": _res(vs.search(q, limit=10)),
            "limit3": _res(vs.search(q, limit=3)),
            "filter_lt3_limit10": _res(vs.search(q, limit=10, filter_metadata={"index": {"$lt": 3}})),
            "filter_lt3_limit2": _res(vs.search(q, limit=2, filter_metadata={"index": {"$lt": 3}})),
            "filter_source_limit4": _res(vs.search(q, limit=4, filter_metadata={"source": "batch_test", "index": {"$gte": 6}})),
            "threshold_09995": _res(vs.search(q, limit=10, threshold=0.9995)),
            "stats_indices": len(vs.get_stats()["indices"]),
        }
    return case


def self_query_case(index_type):
    # README.md:163, examples/basic_usage.py:23-30, tests/test_core.py:135-142
    v = [0.1] * 384
    with tempfile.TemporaryDirectory() as tmp:
        vs = _store(index_type, 384, 1, tmp)
        vs.store("self", v, {"k": "v"})
        return {"index_type": index_type, "dim": 384, "result": _res(vs.search(v, limit=1))}


def random_case(index_type, name, n, dim, shards, k, nq, seed, zero_row=None, dup=None):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim), dtype=np.float32)
    if zero_row is not None:
        X[zero_row] = 0.0
    if dup is not None:
        X[dup[1]] = X[dup[0]]
    Q = np.random.default_rng(seed + 1).standard_normal((nq, dim), dtype=np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        vs = _store(index_type, dim, shards, tmp)
        ids = [f"v{i}" for i in range(n)]
        vs.batch_store({ids[i]: X[i] for i in range(n)},
                       {ids[i]: {"i": i, "even": i % 2 == 0} for i in range(n)})
        placement = [vs._get_shard_for_id(i) for i in ids]
        out = {
            "name": name, "index_type": index_type, "n": n, "dim": dim, "num_shards": shards,
            "k": k, "nq": nq, "seed": seed, "zero_row": zero_row, "dup": dup,
            "placement": placement,
            "results": [_res(vs.search(Q[b], limit=k)) for b in range(nq)],
            "results_filter_even": [_res(vs.search(Q[b], limit=k, filter_metadata={"even": True}))
                                    for b in range(min(nq, 2))],
            "results_big_limit": _res(vs.search(Q[0], limit=n + 7)) if n <= 64 else None,
        }
    for lst in out["results"] + out["results_filter_even"]:
        for r in lst:
            r[2] = None  # metadata is derivable from the id; keep the fixture small
    if out["results_big_limit"]:
        for r in out["results_big_limit"]:
            r[2] = None
    return out


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        os.execve(sys.executable, [sys.executable] + sys.argv, dict(os.environ, PYTHONHASHSEED="0"))
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, REF)
    _install_standins()
    golden = {
        "generator": "tests/golden/make_golden.py (reference VectorStore/FaissIndex/HNSWIndex over exact stand-ins)",
        "ramp": [ramp_case("faiss"), ramp_case("hnsw")],
        "self_query": [self_query_case("faiss"), self_query_case("hnsw")],
        "random": [
            random_case("faiss", "tiny_d5", 37, 5, 3, 4, 3, 11, zero_row=7, dup=(3, 30)),
            random_case("faiss", "small_d4", 200, 4, 2, 10, 4, 12),
            random_case("faiss", "quickstart_c1", 10000, 384, 2, 5, 8, 1234),
            random_case("hnsw", "quickstart_c1_hnsw", 10000, 384, 2, 5, 2, 1234),
        ],
    }
    OUT.write_text(json.dumps(golden, separators=(",", ":")))
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
