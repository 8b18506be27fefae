#!/usr/bin/env python
"""Generate tests/golden/api_probe_golden.json: two scripted call sequences -- one against the reference's own,
unmodified ``VectorStore`` (wdbx/core/vector_store.py), one against its ``WDBX`` facade (wdbx/core/wdbx.py), both
imported from /root/reference and running over the exact faiss stand-in of make_golden.py -- with every call's
outcome (return value, or exception type + message).  Run in the build container only:

    PYTHONHASHSEED=0 python tests/golden/make_api_golden.py

tests replay the same scripts through wdbx_b200 (CPU: numpy engine double; B200: the real engine) and must get the
same outcomes, except at the steps listed in DEVIATIONS, each of which names the documented decision (DESIGN.md
section 1, "decisions where the reference is buggy / undefined") and what ours returns instead."""
import asyncio
import json
import logging
import os
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden  # noqa: E402

OUT = Path(__file__).resolve().parent / "api_probe_golden.json"

# [step name, method, positional args]
STORE_SCRIPT = [
    ["search empty", "search", [[1, 0, 0, 0], 5]],
    ["store a", "store", ["a", [1, 0, 0, 0], {"t": 1}]],
    ["store b", "store", ["b", [0, 1, 0, 0]]],
    ["store not numeric", "store", ["s", "abc"]],
    ["count", "count", []],
    ["batch", "batch_store", [{"c": [1, 1, 0, 0], "d": [0, 0, 1, 0]}, {"c": {"t": 2}}]],
    ["search", "search", [[1, 0.1, 0, 0], 10]],
    ["search limit 0", "search", [[1, 0.1, 0, 0], 0]],
    ["search limit -1", "search", [[1, 0.1, 0, 0], -1]],
    ["search limit 2", "search", [[0.2, 1, 0.1, 0], 2]],
    ["search wrong dim", "search", [[1, 0.1, 0], 3]],
    ["search threshold", "search", [[1, 0, 0, 0], 10, 0.7]],
    ["search threshold exact", "search", [[1, 0, 0, 0], 10, 1.0]],
    ["search negative threshold", "search", [[1, 0.05, 0.01, 0.02], 10, -0.5]],
    ["search filter", "search", [[1, 0, 0, 0], 10, 0.0, {"t": {"$gte": 1}}]],
    ["search filter + threshold", "search", [[1, 0, 0, 0], 10, 0.8, {"t": {"$in": [1, 2]}}]],
    ["search filter none match", "search", [[1, 0, 0, 0], 10, 0.0, {"t": 7}]],
    ["search filter type error", "search", [[1, 0, 0, 0], 10, 0.0, {"t": {"$gt": "x"}}]],
    ["search async", "search_async", [[1, 0.3, 0, 0], 3]],
    ["search async filter", "search_async", [[1, 0.3, 0, 0], 3, 0.0, {"t": 2}]],
    ["get a", "get", ["a"]],
    ["get missing", "get", ["zz"]],
    ["update a", "update_metadata", ["a", {"t": 9}]],
    ["update missing", "update_metadata", ["zz", {"t": 9}]],
    ["get a again", "get", ["a"]],
    ["search sees new metadata", "search", [[1, 0, 0, 0], 1]],
    ["delete missing", "delete", ["zz"]],
    ["delete d", "delete", ["d"]],
    ["delete d again", "delete", ["d"]],
    ["get d", "get", ["d"]],
    ["search after delete", "search", [[0, 0, 1, 0], 10]],
    ["stats keys", "get_stats:keys", []],
    ["stats indices", "get_stats:indices", []],
    ["clear", "clear", []],
    ["count after clear", "count", []],
    ["search cleared", "search", [[1, 0, 0, 0], 3]],
    ["store after clear", "store", ["n", [0, 0, 0, 1], {"k": "v"}]],
    ["search after clear", "search", [[0, 0, 0, 1], 3]],
]
FACADE_SCRIPT = [
    ["initialize", "initialize", []],
    ["store id", "vector_store_async", [[1, 0, 0, 0], {"t": 1}, "a"]],
    ["store id2", "vector_store_async", [[0, 1, 0, 0], None, "b"]],
    ["store wrong dim", "vector_store_async", [[1, 0, 0], {}]],
    ["count", "count_vectors", []],
    ["search", "vector_search", [[1, 0.2, 0, 0], 2]],
    ["search async", "vector_search_async", [[1, 0.2, 0, 0], 2]],
    ["search wrong dim", "vector_search", [[1, 0.2, 0]]],
    ["search async wrong dim", "vector_search_async", [[1, 0.2, 0]]],
    ["search threshold", "vector_search", [[1, 0, 0, 0], 10, 0.5]],
    ["search filter", "vector_search", [[1, 0, 0, 0], 10, 0.0, {"t": 1}]],
    ["get a", "get_vector", ["a"]],
    ["get missing", "get_vector", ["zz"]],
    ["get async", "get_vector_async", ["b"]],
    ["update", "update_metadata", ["b", {"u": 2}]],
    ["update async missing", "update_metadata_async", ["zz", {"u": 2}]],
    ["delete", "delete_vector", ["b"]],
    ["delete async missing", "delete_vector_async", ["zz"]],
    ["stats keys", "get_stats:keys", []],
    ["clear", "clear", []],
    ["count after clear", "count_vectors", []],
    ["shutdown", "shutdown", []],
]
# the operator boundary itself (VectorIndex, indexing.py:18-217): the reference's FaissIndex, ours = the shard facade
# of a one-shard store.  {"__nd__": [...]} stands for a float32 ndarray (the ABC's argument type)
def _v(*x):
    return {"__nd__": list(x)}


INDEX_SCRIPT = [
    ["initialize", "initialize", []],
    ["size empty", "size", []],
    ["search empty", "search", [_v(1, 0, 0, 0), 3]],
    ["add a", "add", ["a", _v(1, 0, 0, 0)]],
    ["add b", "add", ["b", _v(0, 2, 0, 0)]],
    ["add wrong dim", "add", ["w", _v(1, 0, 0)]],
    ["batch_add", "batch_add", [{"c": _v(1, 1, 0, 0), "d": _v(0, 0, 0, 5)}]],
    ["batch_add empty", "batch_add", [{}]],
    ["batch_add ragged", "batch_add", [{"r1": _v(1, 1, 0, 0), "r2": _v(0, 0)}]],
    ["size", "size", []],
    ["search", "search", [_v(1, 0.5, 0, 0), 3]],
    ["search limit > size", "search", [_v(1, 0.5, 0, 0), 50]],
    ["search limit 0", "search", [_v(1, 0.5, 0, 0), 0]],
    ["search wrong dim", "search", [_v(1, 0.5, 0), 2]],
    ["search_async", "search_async", [_v(0, 1, 0, 0), 2]],
    ["remove b", "remove", ["b"]],
    ["remove b again", "remove", ["b"]],
    ["remove_async missing", "remove_async", ["zz"]],
    ["size after remove", "size", []],
    ["optimize", "optimize", []],
    ["stats required keys", "get_stats:required", []],
    ["stats size", "get_stats:size", []],
    ["clear", "clear", []],
    ["size cleared", "size", []],
    ["search cleared", "search", [_v(1, 0, 0, 0), 2]],
    ["add_async after clear", "add_async", ["z", _v(0, 0, 0, 1)]],
    ["search after clear", "search", [_v(0, 0, 0, 1), 2]],
    ["shutdown", "shutdown", []],
]
# step -> [what ours returns instead, the decision behind it]
DEVIATIONS = {
    "store": {
        "search after delete": [["ok", [["a", 0.0, {"t": 9}], ["b", 0.0, {}], ["c", 0.0, {"t": 2}]], "list"],
                                "decision 1 (deleted rows are never returned; the reference returns the row as str(row) "
                                "with {} metadata) and decision 5 (ties: lower insertion id first)"],
        "stats keys": [["ok", ["gpu", "index_type", "indices", "metadata_count", "num_shards", "use_gpu", "vector_count", "vector_dim"], "list"],
                       "additive: the engine's device counters under 'gpu'"],
    },
    "facade": {
        "stats keys": [["ok", ["distributed_enabled", "gpu", "gpu_enabled", "index_type", "indices", "metadata_count", "num_shards",
                               "plugins_enabled", "plugins_loaded", "total_vectors", "use_gpu", "vector_count", "vector_dim",
                               "vector_dimension", "version"], "list"], "additive: 'gpu'"],
    },
}


def norm(x):
    if isinstance(x, float) or type(x).__name__ in ("float32", "float64"):
        return round(float(x), 5)
    if isinstance(x, (list, tuple)):
        return [norm(v) for v in x]
    if isinstance(x, dict):
        return {k: norm(v) for k, v in x.items()}
    return x


def play(obj, script):
    """shared with the tests: run a script against `obj`, normalised outcomes"""
    import numpy as np

    def dec(a):
        if isinstance(a, dict) and "__nd__" in a:
            return np.asarray(a["__nd__"], dtype=np.float32)
        if isinstance(a, dict):
            return {k: dec(v) for k, v in a.items()}
        return a

    loop = asyncio.new_event_loop()
    out = []
    try:
        for name, method, args in script:
            try:
                if method == "get_stats:keys":
                    r = sorted(obj.get_stats().keys())
                elif method == "get_stats:indices":
                    r = len(obj.get_stats()["indices"])
                elif method == "get_stats:required":
                    r = sorted(k for k in obj.get_stats() if k in ("type", "size", "dimension", "gpu_enabled"))
                elif method == "get_stats:size":
                    r = obj.get_stats()["size"]
                else:
                    r = getattr(obj, method)(*[dec(a) for a in args])
                    if asyncio.iscoroutine(r):
                        r = loop.run_until_complete(r)
                out.append(norm(["ok", r]) + [type(r).__name__])
            except Exception as e:   # noqa: BLE001  (type + message are the recorded outcome)
                out.append(["raises", type(e).__name__, str(e)[:60]])
    finally:
        loop.close()
    return out


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        os.execve(sys.executable, [sys.executable] + sys.argv, dict(os.environ, PYTHONHASHSEED="0"))
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, make_golden.REF)
    make_golden._install_standins()
    from wdbx.core.vector_store import VectorStore   # the reference's classes, unmodified
    from wdbx.core.wdbx import WDBX

    from wdbx.core.indexing import FaissIndex

    with tempfile.TemporaryDirectory() as t0:
        index = play(FaissIndex(4, Path(t0) / "ix", None), INDEX_SCRIPT)
    with tempfile.TemporaryDirectory() as t1, tempfile.TemporaryDirectory() as t2:
        store = play(VectorStore(vector_dim=4, data_dir=Path(t1), num_shards=2, index_type="faiss"), STORE_SCRIPT)
        facade = play(WDBX(vector_dimension=4, num_shards=2, data_dir=t2, enable_plugins=False,
                           config={"VECTOR_INDEX_TYPE": "faiss", "INDEX_TYPE": "faiss"}), FACADE_SCRIPT)
    import inspect
    from wdbx.core.indexing import VectorIndex

    def surface(cls):
        out = {}
        for n in dir(cls):
            f = getattr(cls, n)
            if n.startswith("_") or not callable(f):
                continue
            out[n] = {"params": [p for p in inspect.signature(f).parameters if p != "self"],
                      "defaults": {k: v.default for k, v in inspect.signature(f).parameters.items()
                                   if v.default is not inspect.Parameter.empty and isinstance(v.default, (int, float, str, bool, type(None)))},
                      "async": inspect.iscoroutinefunction(f)}
        return out

    surf = {"VectorIndex": surface(VectorIndex), "FaissIndex": surface(FaissIndex), "VectorStore": surface(VectorStore),
            "WDBX": surface(WDBX), "VectorIndex.abstract": sorted(VectorIndex.__abstractmethods__)}
    OUT.write_text(json.dumps({"surface": surf, "generator": "tests/golden/make_api_golden.py (reference VectorStore / WDBX over the exact faiss stand-in)",
                               "index_script": INDEX_SCRIPT, "index": index, "store_script": STORE_SCRIPT, "store": store, "facade_script": FACADE_SCRIPT, "facade": facade,
                               "deviations": DEVIATIONS}, separators=(",", ":")))
    print(f"wrote {OUT}: {len(index)} + {len(store)} + {len(facade)} steps")
    for (n, _, _), o in zip(INDEX_SCRIPT, index):
        print("  index ", n, o)
    for (n, _, _), o in zip(STORE_SCRIPT, store):
        print("  store ", n, o)
    for (n, _, _), o in zip(FACADE_SCRIPT, facade):
        print("  facade", n, o)


if __name__ == "__main__":
    main()
