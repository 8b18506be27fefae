#!/usr/bin/env python
"""Generate tests/golden/loader_golden.json: what the REFERENCE's own, unmodified feeders
(wdbx/utils/data_utils.py: parse_vector :174-231, load_vectors_from_csv :16-108, load_vectors_from_jsonl :111-171,
imported from /root/reference) return -- or raise -- on a grid of inputs and on small CSV / JSONL files (stored
verbatim in the fixture).  Run in the build container only:

    python tests/golden/make_loader_golden.py

tests/test_store_host_logic.py replays everything through wdbx_b200.data_utils."""
import json
import logging
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden  # noqa: E402

OUT = Path(__file__).resolve().parent / "loader_golden.json"

PARSE = [
    [1, 2, 3], [1.5, "2", True], [], ["a"], [[1, 2]], [None],
    "[1, 2, 3]", "[1.5,2.5]", " [1, 2] ", "[]", "[1, 2", "[\"1\", \"2\"]", "[\"a\"]", "[[1, 2]]", "[1, null]", "[1 2 3]", "[1. 2.]",
    "1,2,3", "1, 2 ,3", "1.0e-3,2E5", "1;2", "1,,2", ",", "", "   ", "1 2 3", "1\t2\n3", "1, 2 3", "nan,inf,-inf", "0x10,1",
    "array([1., 2.])", "array([1. 2.])", "array([1, 2, 3])", "np.array([1 2])", "(1 2)", "[1. 2. 3.]\n", "1e400", "١٢", "1_000",
    {"vector": [1, 2]}, {"embedding": "3,4"}, {"values": "[5]"}, {"data": {"vector": "6 7"}}, {"vector": "x", "embedding": [1]},
    {"other": [1]}, {}, 5, 2.5, None, True, (1, 2),
]

CSV_FILES = {
    "named": "id,vec,lang,n\na,\"[1, 2]\",en,1\nb,\"3,4\",de,2\nc,5 6,fr,3\nd,oops,xx,4\ne,\"[7, 8]\",\"multi\r\nline\",5\n",
    "noheader_idx": "r0,\"1,2\",x\nr1,\"3,4\",y\nr2,bad,z\nr3,\"5,6\"\n",
    "semicolon": "id;vec;tag\nq;1,2;t1\nw;[3, 4];t2\n",
    "dup_ids": "id,vec\nk,\"1,2\"\nk,\"3,4\"\n",
    "empty": "",
    "header_only": "id,vec\n",
    "blank_lines": "id,vec\n\na,\"1,2\"\n\nb,\"3,4\"\n",
}
CSV_CALLS = [
    ["named", {"vector_column": "vec", "id_column": "id", "metadata_columns": ["lang", "n", "missing"]}],
    ["named", {"vector_column": "vec"}],
    ["named", {"vector_column": "vec", "id_column": "lang"}],
    ["named", {"vector_column": "nope", "id_column": "id"}],
    ["named", {"vector_column": 1, "id_column": 0, "metadata_columns": [2, 3, 9]}],
    ["named", {"vector_column": 1, "id_column": 0, "skip_header": False}],
    ["named", {"vector_column": 1, "metadata_columns": ["lang"]}],
    ["noheader_idx", {"vector_column": 1, "id_column": 0, "skip_header": False, "metadata_columns": [2]}],
    ["noheader_idx", {"vector_column": 1, "skip_header": True}],
    ["noheader_idx", {"vector_column": 5, "id_column": 0, "skip_header": False}],
    ["semicolon", {"vector_column": "vec", "id_column": "id", "delimiter": ";", "metadata_columns": ["tag"]}],
    ["semicolon", {"vector_column": 1, "id_column": 0, "delimiter": ";"}],
    ["dup_ids", {"vector_column": "vec", "id_column": "id"}],
    ["empty", {"vector_column": "vec"}],
    ["empty", {"vector_column": 0}],
    ["header_only", {"vector_column": "vec", "id_column": "id"}],
    ["blank_lines", {"vector_column": "vec", "id_column": "id"}],
    ["blank_lines", {"vector_column": 1, "id_column": 0}],
    ["missing_file", {"vector_column": "vec"}],
]
JSONL_FILES = {
    "basic": "{\"id\": \"a\", \"emb\": [1, 2], \"lang\": \"en\", \"n\": 1}\n{\"id\": \"b\", \"emb\": \"3,4\", \"lang\": \"de\"}\n"
             "{\"id\": \"c\", \"lang\": \"fr\"}\nnot json\n{\"id\": \"d\", \"emb\": {\"vector\": [5, 6]}}\n\n{\"emb\": [7], \"id\": null}\n[1, 2]\n",
    "noid": "{\"emb\": [1]}\n{\"emb\": [2]}\n",
    "empty": "",
}
JSONL_CALLS = [
    ["basic", {"vector_field": "emb", "id_field": "id"}],
    ["basic", {"vector_field": "emb"}],
    ["basic", {"vector_field": "emb", "id_field": "id", "metadata_fields": ["lang", "zzz"]}],
    ["basic", {"vector_field": "emb", "id_field": "nope"}],
    ["basic", {"vector_field": "vec", "id_field": "id"}],
    ["noid", {"vector_field": "emb"}],
    ["empty", {"vector_field": "emb"}],
    ["missing_file", {"vector_field": "emb"}],
]


def outcome(fn, *a, **kw):
    try:
        return {"ok": fn(*a, **kw)}
    except Exception as e:   # noqa: BLE001  (the type is the recorded outcome)
        return {"raises": type(e).__name__}


def jsonable(x):
    """tuples -> lists, non-string dict keys -> their repr, nan / inf -> strings (JSON has none)"""
    if isinstance(x, dict):
        return {(k if isinstance(k, str) else f"<{k!r}>"): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
        return f"<{x!r}>"
    return x


def main():
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, make_golden.REF)
    make_golden._install_standins()
    from wdbx.utils import data_utils as ref    # the reference's module, unmodified

    golden = {"generator": "tests/golden/make_loader_golden.py (reference wdbx/utils/data_utils.py)",
              "parse_inputs": jsonable(PARSE), "parse": [jsonable(outcome(ref.parse_vector, x)) for x in PARSE],
              "csv_files": CSV_FILES, "csv_calls": CSV_CALLS, "csv": [],
              "jsonl_files": JSONL_FILES, "jsonl_calls": JSONL_CALLS, "jsonl": []}
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in list(CSV_FILES.items()) + list(JSONL_FILES.items()):
            with open(Path(tmp) / name, "w", encoding="utf-8", newline="") as f:   # bytes exactly as in the fixture
                f.write(text)
        for name, kw in CSV_CALLS:
            golden["csv"].append(jsonable(outcome(ref.load_vectors_from_csv, str(Path(tmp) / name), **kw)))
        for name, kw in JSONL_CALLS:
            golden["jsonl"].append(jsonable(outcome(ref.load_vectors_from_jsonl, str(Path(tmp) / name), **kw)))
    OUT.write_text(json.dumps(golden, separators=(",", ":")))
    print(f"wrote {OUT}: {len(PARSE)} parse inputs, {len(CSV_CALLS)} csv calls, {len(JSONL_CALLS)} jsonl calls")


if __name__ == "__main__":
    main()
