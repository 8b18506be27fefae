"""GPU parity of K2b (bf16 tensor-core filter + exact fp32 refine): results must be BIT-IDENTICAL to
the streaming kernel K1 (same ids, same fp32 scores), for every metric, for fp32 and bf16 storage,
including the overflow fallback on adversarial data."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import exact_search as oracle  # noqa: E402


def _engine(dim, nseg=1, dtype="fp32", gemm_min_batch=1, mode=0):
    import wdbx_b200

    os.environ["WDBX_B200_GEMM_MIN_BATCH"] = str(gemm_min_batch)
    os.environ["WDBX_B200_GEMM_MODE"] = str(mode)
    try:
        return wdbx_b200.Engine(device=0, dim=dim, dtype=dtype, num_segments=nseg)
    finally:
        os.environ.pop("WDBX_B200_GEMM_MIN_BATCH", None)
        os.environ.pop("WDBX_B200_GEMM_MODE", None)


def _oracle_check(X, Q, k, metric, scores, gids, counts, dead=None, sample=16):
    for b in np.unique(np.linspace(0, Q.shape[0] - 1, min(sample, Q.shape[0])).astype(int)):
        c = int(counts[b])
        rep = oracle.check_topk(X, Q[b], metric, k, gids[b, :c], scores[b, :c], dead=dead)
        assert rep["count_ok"] and rep["hard_mismatch"] == 0 and rep["recall"] == 1.0, (metric, b, rep)
        assert rep["max_err_over_tol"] <= 1.0 and rep["sorted"], (metric, b, rep)


@pytest.mark.parametrize("n,dim,B,k,metric,dtype", [
    (1000, 64, 128, 10, "ip", "fp32"), (1000, 64, 128, 10, "cosine", "fp32"), (1000, 64, 128, 10, "l2", "fp32"),
    (5000, 768, 200, 10, "cosine", "fp32"), (20011, 384, 64, 5, "cosine", "fp32"), (3000, 1536, 130, 10, "l2", "fp32"),
    (777, 100, 33, 32, "ip", "fp32"), (70000, 96, 512, 10, "cosine", "fp32"), (300, 20, 256, 10, "cosine", "fp32"),
    (30000, 384, 96, 20, "ip", "bf16"), (9000, 200, 70, 10, "cosine", "bf16"), (150000, 128, 300, 10, "l2", "fp32"),
    # small batches: several refine CTAs per query + K3 merge of their partial lists
    (200000, 256, 1, 10, "cosine", "fp32"), (50000, 768, 3, 16, "l2", "fp32"), (120000, 128, 8, 10, "ip", "fp32"),
    (40000, 384, 2, 1, "cosine", "bf16"), (60000, 256, 4, 25, "cosine", "fp32"), (20000, 96, 300, 32, "l2", "fp32")])
def test_filter_refine_is_bit_identical_to_scan(built_lib, n, dim, B, k, metric, dtype):
    rng = np.random.default_rng(n + dim + B)
    X = rng.standard_normal((n, dim), dtype=np.float32)
    Q = rng.standard_normal((B, dim), dtype=np.float32)
    e2, e1 = _engine(dim, dtype=dtype, gemm_min_batch=1), _engine(dim, dtype=dtype, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    s2, g2, c2 = e2.search_host(Q, k, metric=metric)     # K2b
    s1, g1, c1 = e1.search_host(Q, k, metric=metric)     # K1
    np.testing.assert_array_equal(c2, c1)
    np.testing.assert_array_equal(g2, g1)
    np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))   # bit-identical scores
    Xs = oracle.bf16_round(X) if dtype == "bf16" else X
    _oracle_check(Xs, Q, k, metric, s2, g2, c2)
    e2.close(); e1.close()


@pytest.mark.parametrize("B", [160, 6])   # 128-query kernel (CTA pair) / small-batch kernel
def test_filter_segments_tombstones_edge_rows(built_lib, B):
    rng = np.random.default_rng(78)
    n, dim, k = 9000, 384, 10
    X = rng.standard_normal((n, dim), dtype=np.float32)
    X[100] = 0.0
    X[4000] = X[17]
    X[6000, 5] = np.nan
    Q = rng.standard_normal((B, dim), dtype=np.float32)
    Q[3] = X[17]
    Q[5] = 0.0
    seg_of = rng.integers(0, 3, size=n)
    dead = np.zeros(n, bool)
    dead[[5, 17, 8000]] = True
    engs = [_engine(dim, nseg=3, gemm_min_batch=1), _engine(dim, nseg=3, gemm_min_batch=0)]
    for eng in engs:
        for s_ in range(3):
            rows = np.flatnonzero(seg_of == s_)
            eng.append(s_, X[rows], gids=rows.astype(np.uint32))
        for r in np.flatnonzero(dead):
            seg = int(seg_of[r])
            eng.tombstone(seg, int(np.sum(seg_of[:r] == seg)))
    for metric in ("cosine", "ip", "l2"):
        sg, gg, cg = engs[0].search_host(Q, k, metric=metric)
        ss, gs, cs = engs[1].search_host(Q, k, metric=metric)
        np.testing.assert_array_equal(gg, gs)
        np.testing.assert_array_equal(sg.view(np.uint32), ss.view(np.uint32))
    assert gg[3, 0] != 17
    # overwrite + append after the shadow exists, then search again
    for eng in engs:
        eng.overwrite(0, 3, X[50] * 3.0)
        eng.append(1, X[:500] + 1.0, gids=np.arange(20000, 20500, dtype=np.uint32))
    sg, gg, cg = engs[0].search_host(Q, k)
    ss, gs, cs = engs[1].search_host(Q, k)
    np.testing.assert_array_equal(gg, gs)
    np.testing.assert_array_equal(sg.view(np.uint32), ss.view(np.uint32))
    for eng in engs:
        eng.close()


def test_filter_overflow_falls_back_to_exact_scan(built_lib):
    """Adversarial: thousands of (near-)identical rows make every row a candidate -> the candidate list
    overflows -> flagged queries are re-run by K1; results stay exact."""
    rng = np.random.default_rng(3)
    n, dim, B, k = 12000, 64, 64, 10
    base = rng.standard_normal(dim).astype(np.float32)
    X = np.tile(base, (n, 1)) + rng.standard_normal((n, dim)).astype(np.float32) * 1e-4
    Q = np.tile(base, (B, 1)) + rng.standard_normal((B, dim)).astype(np.float32) * 1e-3
    Q[1] = rng.standard_normal(dim)          # an ordinary query in the same batch
    e2, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    s2, g2, c2 = e2.search_host(Q, k)
    s1, g1, c1 = e1.search_host(Q, k)
    np.testing.assert_array_equal(g2, g1)
    np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
    assert np.all(c2 == k)
    e2.close(); e1.close()


def test_small_k_larger_than_rows_and_empty(built_lib):
    dim = 32
    e2 = _engine(dim, gemm_min_batch=1)
    Q = np.random.default_rng(0).standard_normal((50, dim)).astype(np.float32)
    s, g, c = e2.search_host(Q, 10)
    assert np.all(c == 0) and np.all(g == -1)
    X = np.random.default_rng(1).standard_normal((7, dim)).astype(np.float32)
    e2.append(0, X)
    s, g, c = e2.search_host(Q, 10)
    assert np.all(c == 7) and np.all(g[:, 7:] == -1)
    _oracle_check(X, Q, 10, "cosine", s, g, c)
    e2.close()


def test_small_batches_route_by_size_and_report_the_dominant_kernel(built_lib):
    """fp32 segments above the size threshold serve even single queries through the bf16-shadow filter
    (half the HBM bytes, bit-identical results); wdbx_b200_set_kernel_timing reports which kernel ran."""
    import wdbx_b200

    rng = np.random.default_rng(5)
    X = rng.standard_normal((50000, 128), dtype=np.float32)
    q = rng.standard_normal((1, 128), dtype=np.float32)
    got = {}
    for name, mb in (("filter", "0"), ("scan", "-1")):
        os.environ["WDBX_B200_SHADOW_MIN_MB"] = mb
        try:
            e = wdbx_b200.Engine(device=0, dim=128, dtype="fp32", num_segments=1)
        finally:
            os.environ.pop("WDBX_B200_SHADOW_MIN_MB", None)
        e.append(0, X)
        e.set_kernel_timing(True)
        got[name] = e.search_host(q, 10, metric="cosine")
        st = e.stats()
        assert st["last_kernel"] == (2 if name == "filter" else 1) and st["last_kernel_ms"] > 0.0
        e.set_kernel_timing(False)
        e.close()
    for a, b in zip(got["filter"], got["scan"]):
        np.testing.assert_array_equal(np.asarray(a).view(np.uint32) if a.dtype == np.float32 else a,
                                      np.asarray(b).view(np.uint32) if b.dtype == np.float32 else b)


@pytest.mark.parametrize("B", [1, 20])
def test_ascending_scores_worst_case_for_a_running_bound(built_lib, B):
    """Rows arrive in ascending score order: every row beats the bound known when it is seen, so (almost)
    every row is a candidate -- the regions overflow and the flagged queries are re-run by the exact scan.
    The answer must not change."""
    rng = np.random.default_rng(99)
    n, dim, k = 60000, 64, 10
    u = rng.standard_normal(dim).astype(np.float32)
    u /= np.linalg.norm(u)
    ramp = (np.arange(1, n + 1, dtype=np.float32) / n)[:, None]
    X = (ramp * u[None, :] + 1e-4 * rng.standard_normal((n, dim), dtype=np.float32)).astype(np.float32)
    Q = (u[None, :] + 0.01 * rng.standard_normal((B, dim), dtype=np.float32)).astype(np.float32)
    e2, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    for metric in ("ip", "l2"):
        s2, g2, c2 = e2.search_host(Q, k, metric=metric)
        s1, g1, c1 = e1.search_host(Q, k, metric=metric)
        np.testing.assert_array_equal(c2, c1)
        np.testing.assert_array_equal(g2, g1)
        np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
        _oracle_check(X, Q, k, metric, s2, g2, c2, sample=3)
    e2.close(); e1.close()


@pytest.mark.parametrize("B", [3, 40])
def test_zero_query_and_zero_rows_on_the_filter_path(built_lib, B):
    """A zero query scores every row 0 (cosine / ip): the bound never separates anything, every row is a
    candidate, and the answer is the k lowest gids -- exactly what the scan returns."""
    rng = np.random.default_rng(17)
    n, dim, k = 5000, 32, 10
    X = rng.standard_normal((n, dim), dtype=np.float32)
    X[11] = 0.0
    Q = rng.standard_normal((B, dim), dtype=np.float32)
    Q[1] = 0.0
    e2, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    for metric in ("cosine", "ip", "l2"):
        s2, g2, c2 = e2.search_host(Q, k, metric=metric)
        s1, g1, c1 = e1.search_host(Q, k, metric=metric)
        np.testing.assert_array_equal(c2, c1)
        np.testing.assert_array_equal(g2, g1)
        np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
        if metric != "l2":
            assert g2[1].tolist() == list(range(k)) and not s2[1].any()
    e2.close(); e1.close()
