"""Checks shared by the CPU (numpy double) and GPU (real engine) suites: replay the reference's
golden outputs (tests/golden/reference_golden.json) through wdbx_b200.VectorStore."""
import numpy as np

TIE = 2e-6


def assert_same_results(got, want, what=""):
    """got/want: [(id, score, meta-or-None)].  Same length, scores within 1e-5 rel + 1e-6, same ids in
    the same order except inside exact-tie groups (the reference's tie order is library-defined)."""
    assert len(got) == len(want), (what, len(got), len(want), got, want)
    for j, (g, w) in enumerate(zip(got, want)):
        assert abs(g[1] - w[1]) <= 1e-5 * abs(w[1]) + 1e-6, (what, j, g, w)
    i = 0
    while i < len(want):
        j = i
        while j + 1 < len(want) and abs(want[j + 1][1] - want[i][1]) <= TIE:
            j += 1
        assert sorted(x[0] for x in got[i:j + 1]) == sorted(x[0] for x in want[i:j + 1]), (what, i, got, want)
        i = j + 1
    for g, w in zip(got, want):
        if w[2] is not None and g[0] == w[0]:
            assert g[2] == w[2], (what, g, w)


def load_with_placement(store, ids, vectors, placement, metadata):
    """Put every id on the shard the reference chose (its hash placement is process-salted,
    vector_store.py:188-190), through the per-shard index facades (the operator boundary)."""
    for s in range(store.num_shards):
        sel = [i for i, vid in enumerate(ids) if placement[vid] == s]
        if sel:
            assert store.indices[s].batch_add({ids[i]: np.asarray(vectors[i], np.float32) for i in sel})
    for vid in ids:
        store.metadata[vid] = metadata.get(vid, {})


def check_ramp(make_store, case):
    store = make_store(case["dim"], case["num_shards"])
    ids = list(case["vectors"].keys())
    load_with_placement(store, ids, [case["vectors"][i] for i in ids], case["placement"], case["metadata"])
    q = case["query"]
    assert store.count() == 10
    assert_same_results(store.search(q, limit=1), case["limit1"], "limit1")
    assert store.search(q, limit=1)[0][0] == "vec_5"                      # tests/test_core.py:218-222
    assert_same_results(store.search(q, limit=10), case["limit10"], "limit10")
    assert_same_results(store.search(q, limit=3), case["limit3"], "limit3")
    f = store.search(q, limit=10, filter_metadata={"index": {"$lt": 3}})
    assert len(f) == 3 and all(m["index"] < 3 for _, _, m in f)             # tests/test_core.py:225-230
    assert_same_results(f, case["filter_lt3_limit10"], "filter10")
    assert_same_results(store.search(q, limit=2, filter_metadata={"index": {"$lt": 3}}),
                        case["filter_lt3_limit2"], "filter2 (post-filter truncation quirk)")
    assert_same_results(store.search(q, limit=4, filter_metadata={"source": "batch_test", "index": {"$gte": 6}}),
                        case["filter_source_limit4"], "filter4")
    assert_same_results(store.search(q, limit=10, threshold=0.9995), case["threshold_09995"], "threshold")
    assert len(store.get_stats()["indices"]) == case["stats_indices"] == 2  # tests/test_core.py:341
    store.close()


def check_self_query(make_store, case):
    store = make_store(case["dim"], 1)
    v = [0.1] * case["dim"]
    assert store.store("self", v, {"k": "v"})
    got = store.search(v, limit=1)
    assert got[0][0] == "self" and got[0][1] > 0.99                         # tests/test_core.py:135-142
    assert_same_results(got, case["result"], "self")
    store.close()


def random_inputs(case):
    rng = np.random.default_rng(case["seed"])
    X = rng.standard_normal((case["n"], case["dim"]), dtype=np.float32)
    if case["zero_row"] is not None:
        X[case["zero_row"]] = 0.0
    if case["dup"] is not None:
        X[case["dup"][1]] = X[case["dup"][0]]
    Q = np.random.default_rng(case["seed"] + 1).standard_normal((case["nq"], case["dim"]), dtype=np.float32)
    return X, Q


def check_random(make_store, case):
    X, Q = random_inputs(case)
    n = case["n"]
    store = make_store(case["dim"], case["num_shards"])
    ids = [f"v{i}" for i in range(n)]
    placement = {ids[i]: case["placement"][i] for i in range(n)}
    meta = {ids[i]: {"i": i, "even": i % 2 == 0} for i in range(n)}
    load_with_placement(store, ids, X, placement, meta)
    for b in range(case["nq"]):
        assert_same_results(store.search(Q[b].tolist(), limit=case["k"]), case["results"][b], f"{case['name']} q{b}")
    for b, want in enumerate(case["results_filter_even"]):
        assert_same_results(store.search(Q[b].tolist(), limit=case["k"], filter_metadata={"even": True}), want,
                            f"{case['name']} filter q{b}")
    if case["results_big_limit"]:
        assert_same_results(store.search(Q[0].tolist(), limit=n + 7), case["results_big_limit"], "big limit")
    # per-shard operator call == reference FaissIndex.search on that shard
    r0 = store.indices[0].search(Q[0], limit=case["k"])
    want0 = [w for w in sorted(
        ((vid, sc) for vid, sc, _ in store.search(Q[0].tolist(), limit=n)
         if placement[vid] == 0), key=lambda t: -t[1])][: case["k"]]
    assert [r[0] for r in r0] == [w[0] for w in want0]
    store.close()
