#!/usr/bin/env python
"""Generate tests/golden/filter_ops_golden.json: the outcome of the REFERENCE's own, unmodified
``VectorStore._matches_filter`` (wdbx/core/vector_store.py:414-463, imported from /root/reference) on a grid of
(metadata, filter) pairs -- every operator of its Mongo-style ladder, missing keys, type mismatches (which raise),
unknown operators, several operators in one clause (only the first is looked at), empty clauses.  Run in the build
container only:

    python tests/golden/make_filter_golden.py

tests/test_store_host_logic.py replays the grid through wdbx_b200.VectorStore._matches_filter."""
import json
import logging
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden  # noqa: E402  (stand-ins for the absent faiss / hnswlib wheels)

OUT = Path(__file__).resolve().parent / "filter_ops_golden.json"

METADATA = [
    {}, {"a": 1}, {"a": 1.5}, {"a": 0}, {"a": -2}, {"a": "x"}, {"a": ""}, {"a": None}, {"a": True}, {"a": False},
    {"a": [1, 2]}, {"a": {"x": 1}}, {"b": 2}, {"a": 3, "b": "y"}, {"a": 1, "b": 2, "c": "z"},
]
FILTERS = [
    {}, {"a": 1}, {"a": 1.0}, {"a": "x"}, {"a": None}, {"a": True}, {"a": [1, 2]}, {"a": {"x": 1}}, {"b": 2}, {"a": 1, "b": 2},
    {"a": {"$gt": 1}}, {"a": {"$gt": 0}}, {"a": {"$gt": "a"}}, {"a": {"$lt": 1}}, {"a": {"$lt": 2}}, {"a": {"$lt": "y"}},
    {"a": {"$gte": 1}}, {"a": {"$gte": 1.5}}, {"a": {"$lte": 1}}, {"a": {"$lte": 0}}, {"a": {"$lte": None}},
    {"a": {"$in": [1, "x"]}}, {"a": {"$in": []}}, {"a": {"$in": "xyz"}}, {"a": {"$in": [None, True]}}, {"a": {"$in": 5}},
    {"a": {"$nin": [1]}}, {"a": {"$nin": []}}, {"a": {"$nin": ["x", 3]}}, {"a": {"$nin": 5}},
    {"a": {"$exists": True}}, {"a": {"$exists": False}}, {"a": {"$exists": 1}}, {"a": {"$exists": 0}}, {"a": {"$exists": None}},
    {"b": {"$exists": True}}, {"c": {"$exists": False}},
    {"a": {"$unknown": 1}}, {"a": {"$gt": 0, "$lt": 2}}, {"a": {"$lt": 2, "$gt": 5}}, {"a": {}},
    {"a": {"$gte": 1}, "b": {"$in": [2, "y"]}}, {"a": {"$gt": 0}, "c": "z"}, {"b": {"$lt": 3}, "a": {"$nin": [3]}},
]


def main():
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, make_golden.REF)
    make_golden._install_standins()
    from wdbx.core.vector_store import VectorStore   # the reference's class, unmodified

    class Holder:                                    # _matches_filter only touches self.metadata
        def __init__(self, md):
            self.metadata = {"id": md}

    rows = []
    for md in METADATA:
        for f in FILTERS:
            try:
                out = bool(VectorStore._matches_filter(Holder(md), "id", f))
            except Exception as e:                    # noqa: BLE001  (the type is the recorded outcome)
                out = "raises:" + type(e).__name__
            rows.append(out)
    # a vector without any metadata entry at all
    missing = []
    for f in FILTERS:
        try:
            h = Holder({})
            h.metadata = {}
            missing.append(bool(VectorStore._matches_filter(h, "id", f)))
        except Exception as e:                        # noqa: BLE001
            missing.append("raises:" + type(e).__name__)
    OUT.write_text(json.dumps({"generator": "tests/golden/make_filter_golden.py (reference VectorStore._matches_filter)",
                               "metadata": METADATA, "filters": FILTERS, "outcomes": rows, "no_entry": missing},
                              separators=(",", ":")))
    print(f"wrote {OUT}: {len(rows)} pairs, {sum(1 for r in rows if r is True)} match, "
          f"{sum(1 for r in rows if isinstance(r, str))} raise")


if __name__ == "__main__":
    main()
