"""CPU: pin the oracle (numpy + C restatement) against the reference's own outputs."""
import ctypes as C

import numpy as np
import pytest

from oracle import exact_search as oracle
from tests.golden_checks import assert_same_results, random_inputs


def _oracle_store(case_dim, shards, metric="cosine"):
    return oracle.OracleStore(case_dim, shards, metric)


def test_ramp_fixture_matches_reference(golden):
    for case in golden["ramp"]:
        st = _oracle_store(4, 2)
        for vid, vec in case["vectors"].items():
            st.add(case["placement"][vid], vid, vec, case["metadata"][vid])
        q = case["query"]
        assert_same_results(st.search(q, 1), case["limit1"])
        assert_same_results(st.search(q, 10), case["limit10"])
        assert_same_results(st.search(q, 10, filter_metadata={"index": {"$lt": 3}}), case["filter_lt3_limit10"])
        assert_same_results(st.search(q, 2, filter_metadata={"index": {"$lt": 3}}), case["filter_lt3_limit2"])
        assert_same_results(st.search(q, 10, threshold=0.9995), case["threshold_09995"])


def test_survey_golden_scores(golden):
    """SURVEY.md section 8c lists the ramp scores to 9 digits."""
    want = [("vec_5", 1.0), ("vec_6", 0.999750078), ("vec_4", 0.999543846), ("vec_7", 0.999217749),
            ("vec_8", 0.998585820), ("vec_9", 0.997936130), ("vec_3", 0.997323334), ("vec_2", 0.990375102),
            ("vec_1", 0.968863964), ("vec_0", 0.891484976)]
    got = golden["ramp"][0]["limit10"]
    for (wid, ws), (gid, gs, _) in zip(want, got):
        assert wid == gid and abs(ws - gs) < 5e-8


def test_self_query(golden):
    for case in golden["self_query"]:
        st = _oracle_store(384, 1)
        st.add(0, "self", [0.1] * 384, {"k": "v"})
        assert_same_results(st.search([0.1] * 384, 1), case["result"])


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_random_cases(golden, idx):
    case = golden["random"][idx]
    X, Q = random_inputs(case)
    st = _oracle_store(case["dim"], case["num_shards"])
    for i in range(case["n"]):
        st.add(case["placement"][i], f"v{i}", X[i], {"i": i, "even": i % 2 == 0})
    for b in range(case["nq"]):
        assert_same_results(st.search(Q[b], case["k"]), case["results"][b])
    for b, want in enumerate(case["results_filter_even"]):
        assert_same_results(st.search(Q[b], case["k"], filter_metadata={"even": True}), want)
    if case["results_big_limit"]:
        assert_same_results(st.search(Q[0], case["n"] + 7), case["results_big_limit"])


def test_c_restatement_matches_numpy_and_golden(golden):
    import __graft_entry__ as ge

    lib = C.CDLL(str(ge.build_oracle()))
    lib.oracle_flat_search.restype = C.c_int
    lib.oracle_flat_search.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_normalize_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    case = golden["random"][2]  # quick-start C1: 10k x 384
    X, Q = random_inputs(case)
    Xn = np.ascontiguousarray(X.copy())
    lib.oracle_normalize_rows(Xn.ctypes.data, Xn.shape[0], Xn.shape[1])
    np.testing.assert_allclose(Xn, oracle.normalize_rows(X), rtol=2e-6, atol=1e-7)
    k = case["k"]
    for b in range(case["nq"]):
        qn = oracle.normalize(Q[b])
        rows = np.empty(k, np.int64)
        sc = np.empty(k, np.float32)
        n = lib.oracle_flat_search(Xn.ctypes.data, Xn.shape[0], Xn.shape[1], qn.ctypes.data, 0, k, None,
                                   rows.ctypes.data, sc.ctypes.data, 0)
        assert n == k
        want = case["results"][b]
        assert [f"v{r}" for r in rows] == [w[0] for w in want]
        np.testing.assert_allclose(sc, [w[1] for w in want], rtol=1e-5, atol=1e-6)
    # l2 / dead mask / k > n
    rows = np.empty(50, np.int64)
    sc = np.empty(50, np.float32)
    small = np.ascontiguousarray(X[:20])
    dead = np.zeros(20, np.uint8)
    dead[4] = 1
    n = lib.oracle_flat_search(small.ctypes.data, 20, X.shape[1], Q[0].ctypes.data, 2, 50, dead.ctypes.data,
                               rows.ctypes.data, sc.ctypes.data, 3)
    assert n == 19
    want_rows, want_s = oracle.topk_desc(oracle.scores_fp32(small, Q[0], "l2"), 50, dead=dead.astype(bool))
    assert list(rows[:n]) == list(want_rows)
    np.testing.assert_allclose(sc[:n], want_s, rtol=1e-5)


def test_topk_tie_rule_and_nan():
    s = np.array([1.0, 3.0, 3.0, np.nan, -np.inf, 2.0], np.float32)
    rows, val = oracle.topk_desc(s, 6)
    assert list(rows) == [1, 2, 5, 0, 3, 4]
    rows, _ = oracle.topk_desc(s, 2)
    assert list(rows) == [1, 2]


def test_bf16_round_matches_torch():
    import torch

    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 100
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(oracle.bf16_round(x), want)
