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
    (40000, 384, 2, 1, "cosine", "bf16"), (60000, 256, 4, 25, "cosine", "fp32"), (20000, 96, 300, 32, "l2", "fp32"),
    # 32 < k <= 128: four keys per lane in every register list, shared lower-bound list of up to 128 slots
    (90000, 384, 1, 100, "cosine", "fp32"), (50000, 256, 5, 128, "ip", "fp32"), (30000, 768, 12, 33, "l2", "fp32"),
    (40000, 384, 40, 100, "cosine", "fp32"), (25000, 128, 200, 64, "l2", "fp32"), (30000, 384, 3, 100, "ip", "bf16"),
    (60, 32, 2, 100, "cosine", "fp32"), (20000, 384, 160, 128, "ip", "bf16")])
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
        assert (st["last_kernel"] in (2, 3) if name == "filter" else st["last_kernel"] == 1) and st["last_kernel_ms"] > 0.0
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


# ----------------------------------------------------------------------------- rigor of the filter's error bound
# bf16 has 8 significand bits: RNE moves an operand by up to 2^-8 relative, and when every element of a row
# rounds the SAME way the errors add up coherently instead of averaging out.  The filter's bound is derived
# from the actual rounding residuals (|x - bf16(x)| per row, |q - bf16(q)| per query), so such data only
# widens eps; a bound that assumed 2^-9 per operand (round 1) dropped the true top-1 of the first case below.
def _judge_counter_example(n_fill, dim_fill_seed=0):
    """D = 608.  q = 301 x 1.0039 ++ 300 x 1.00391 ++ 0;  A = 301 x 1.0039 ++ 0 (every element rounds DOWN to 1.0);
    B = 0 x 301 ++ 300 x 1.00391 ++ 0 (every element rounds UP to 1.0078125).  Exact cosines: A 0.70769 > B 0.70652,
    bf16 filter scores: A 0.70220 < B 0.71203.  B sits at row 0, A at row 1 (same warp, same tile)."""
    D = 608
    q = np.zeros(D, np.float32)
    q[:301] = 1.0039
    q[301:601] = 1.00391
    A = np.zeros(D, np.float32)
    A[:301] = 1.0039
    Bv = np.zeros(D, np.float32)
    Bv[301:601] = 1.00391
    rng = np.random.default_rng(dim_fill_seed)
    fill = rng.standard_normal((n_fill, D)).astype(np.float32)      # cosine ~ 0 +- 0.04 against q
    X = np.concatenate([Bv[None], A[None], fill]).astype(np.float32)
    return X, q


@pytest.mark.parametrize("B", [1, 40, 160])      # small-batch kernel / 128-query kernel / CTA-pair kernel
@pytest.mark.parametrize("metric", ["cosine", "ip"])
def test_bound_is_rigorous(built_lib, B, metric):
    X, q = _judge_counter_example(3000)
    Q = np.tile(q, (B, 1))
    if B > 1:   # the other queries of the batch are small perturbations: same adversarial structure
        Q[1:] *= (1.0 + 1e-3 * np.arange(1, B, dtype=np.float32))[:, None]
    e2, e1 = _engine(X.shape[1], gemm_min_batch=1), _engine(X.shape[1], gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    for k in (1, 2, 10):
        s2, g2, c2 = e2.search_host(Q, k, metric=metric)
        s1, g1, c1 = e1.search_host(Q, k, metric=metric)
        assert g1[0, 0] == 1                         # the exact scan ranks A first
        np.testing.assert_array_equal(g2, g1)        # ... and so does the filter path
        np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
        _oracle_check(X, Q, k, metric, s2, g2, c2, sample=3)
    e2.close(); e1.close()


@pytest.mark.parametrize("B", [1, 40])
def test_bound_is_rigorous_bf16_store(built_lib, B):
    """bf16 storage: the rows are exact, only the query is rounded.  q's elements sit just above bf16 midpoints in
    two groups that round in opposite directions; A and B are bf16-exact indicator rows of the two groups."""
    D = 608
    q = np.zeros(D, np.float32)
    q[:301] = 1.0039
    q[301:601] = 1.00391
    A = np.zeros(D, np.float32)
    A[:301] = 1.0
    Bv = np.zeros(D, np.float32)
    Bv[301:601] = 1.0
    rng = np.random.default_rng(4)
    fill = oracle.bf16_round(rng.standard_normal((3000, D)).astype(np.float32))
    X = np.concatenate([Bv[None], A[None], fill]).astype(np.float32)
    Q = np.tile(q, (B, 1))
    e2, e1 = _engine(D, dtype="bf16", gemm_min_batch=1), _engine(D, dtype="bf16", gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    for metric in ("cosine", "ip"):
        s2, g2, c2 = e2.search_host(Q, 1, metric=metric)
        s1, g1, c1 = e1.search_host(Q, 1, metric=metric)
        assert g1[0, 0] == 1
        np.testing.assert_array_equal(g2, g1)
        np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
    e2.close(); e1.close()


@pytest.mark.parametrize("seed", range(6))
def test_few_level_coherent_rounding_sweep(built_lib, seed):
    """Quantised / few-level embeddings: every stored value is one of a handful of levels that sit just
    above or below bf16 midpoints, all with the same sign, on random supports -- the shape of data on which
    per-element rounding errors do not average out.  Filter path == exact scan, bit for bit, for all metrics."""
    rng = np.random.default_rng(1000 + seed)
    dim = (64, 200, 608)[seed % 3]
    n, B, k = 5000, (1, 7, 40, 160, 1, 33)[seed], (1, 10, 5, 10, 32, 3)[seed]
    levels = np.array([1.0039, 1.00391, 0.50195, 0.50196, 2.0078, 0.25098, 1.9922, 0.99609], np.float32)
    lv = levels[rng.integers(0, len(levels), size=n)]                       # one level per row
    support = rng.random((n, dim)) < rng.uniform(0.2, 0.9, size=(n, 1))     # random support per row
    X = (support * lv[:, None]).astype(np.float32)
    X[rng.integers(0, n, 20)] *= -1.0                                       # a few negated rows
    ql = levels[rng.integers(0, len(levels), size=(B, 1))]
    Q = ((rng.random((B, dim)) < 0.7) * ql).astype(np.float32)
    Q[:, : dim // 2] *= (1.0 + 2.0 ** -9)                                   # two groups rounding differently
    e2, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    for metric in ("cosine", "ip", "l2"):
        s2, g2, c2 = e2.search_host(Q, k, metric=metric)
        s1, g1, c1 = e1.search_host(Q, k, metric=metric)
        np.testing.assert_array_equal(c2, c1)
        np.testing.assert_array_equal(g2, g1)
        np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
    e2.close(); e1.close()


def test_l2_bound_covers_norm_rounding_for_large_rows(built_lib):
    """l2 via the expanded form: when |x| >> |q| the fp32 rounding of the stored |x|^2 dwarfs the bf16 term.
    Near-duplicate large rows whose distances to q differ only in the cross term must still be ranked exactly."""
    rng = np.random.default_rng(31)
    dim, n, k = 256, 4000, 5
    base = (rng.standard_normal(dim) * 300.0).astype(np.float32)
    X = (base[None, :] + rng.standard_normal((n, dim)).astype(np.float32) * 0.05).astype(np.float32)
    Q = (base[None, :] * 1e-3 + rng.standard_normal((3, dim)).astype(np.float32) * 0.1).astype(np.float32)
    e2, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
    s2, g2, c2 = e2.search_host(Q, k, metric="l2")
    s1, g1, c1 = e1.search_host(Q, k, metric="l2")
    np.testing.assert_array_equal(g2, g1)
    np.testing.assert_array_equal(s2.view(np.uint32), s1.view(np.uint32))
    e2.close(); e1.close()


@pytest.mark.parametrize("B,k", [(1, 10), (5, 10), (2, 100)])
def test_overlapping_searches_keep_results_and_order(built_lib, B, k):
    """Opt-in `overlap`: consecutive device-resident searches on one stream overlap on the device (double-buffered
    per-search state, search-number ordering).  Every one of 300 back-to-back searches -- distinct queries, own output
    buffers, a streaming-scan search (k = 200) thrown in every 50 -- must equal the fully ordered run bit for bit."""
    import torch
    from wdbx_b200.engine import new_out

    n, dim = 400_000, 256
    dev = torch.device("cuda", 0)
    os.environ["WDBX_B200_SHADOW_MIN_MB"] = "0"
    try:
        eng = _engine(dim)
    finally:
        os.environ.pop("WDBX_B200_SHADOW_MIN_MB", None)
    g = torch.Generator(device=dev).manual_seed(3)
    X = torch.randn((n, dim), generator=g, device=dev)
    X[1000:1040] = X[7]                      # a clump of duplicates: ties by gid
    eng.append(0, X)
    Q = torch.randn((300, B, dim), generator=g, device=dev)
    Q[17, 0] = X[7]

    def run(overlap):
        eng.set_option("overlap", 1 if overlap else 0)
        outs = [new_out(B, k, dev) for _ in range(300)]
        big = []
        torch.cuda.synchronize()
        for i in range(300):
            eng.search(Q[i], k, "cosine", out=outs[i])
            if i % 50 == 49:
                big.append(eng.search(Q[i][:1], 200, "cosine"))      # K1 key dump + select between fused searches
        torch.cuda.synchronize()
        return [o["keys"].clone() for o in outs], [b["keys"].clone() for b in big], [o["counts"].clone() for o in outs]

    ref_keys, ref_big, ref_cnt = run(False)
    for _ in range(2):
        got_keys, got_big, got_cnt = run(True)
        for i in range(300):
            assert torch.equal(got_keys[i], ref_keys[i]), i
            assert torch.equal(got_cnt[i], ref_cnt[i]), i
        for a, b in zip(got_big, ref_big):
            assert torch.equal(a, b)
    # the fully ordered run itself is right
    s = (X @ Q[17, 0]) / (X.norm(dim=1) * Q[17, 0].norm())
    want = torch.topk(s, k).indices
    gids = (~(ref_keys[17][0] & 0xFFFFFFFF)) & 0xFFFFFFFF
    assert set(gids.tolist()) == set(want.tolist()) or k > 40
    eng.set_option("overlap", 0)
    eng.close()


@pytest.mark.parametrize("metric", ["cosine", "ip", "l2"])
@pytest.mark.parametrize("kind", ["outliers", "cauchy", "few_level", "tiny_and_huge", "near_ties"])
def test_int8_shadow_bound_is_rigorous_on_hostile_rows(built_lib, metric, kind):
    """The small-batch kernel streams an int8 shadow (x ~ sx * xi, symmetric per-row scale).  Its error bound uses the
    ACTUAL residual norm of every row, so rows that quantise badly -- one huge element eating the scale, heavy tails,
    few-level data sitting between the int8 grid points, rows spanning 12 orders of magnitude, thousands of
    near-ties -- only make the filter pass more candidates; the answer must stay bit-identical to the streaming scan."""
    rng = np.random.default_rng(hash((metric, kind)) % (2 ** 31))
    n, dim = 60000, 320
    if kind == "outliers":
        X = rng.standard_normal((n, dim)).astype(np.float32)
        X[np.arange(n), rng.integers(0, dim, n)] *= rng.choice([50.0, 500.0, 5000.0], n).astype(np.float32)
    elif kind == "cauchy":
        X = rng.standard_cauchy((n, dim)).astype(np.float32)
    elif kind == "few_level":
        X = rng.choice(np.array([-1.0, -0.51, 0.0, 0.49, 1.0], np.float32) * 1.00393, (n, dim))
    elif kind == "tiny_and_huge":
        X = rng.standard_normal((n, dim)).astype(np.float32) * (10.0 ** rng.integers(-6, 6, (n, 1))).astype(np.float32)
    else:
        base = rng.standard_normal(dim).astype(np.float32)
        X = base[None, :] + 1e-3 * rng.standard_normal((n, dim)).astype(np.float32)
    Q = rng.standard_normal((5, dim)).astype(np.float32)
    Q[1] = X[123]
    Q[2, rng.integers(0, dim)] = 300.0                      # a query that quantises badly too
    k = 10
    e8, e16, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    e16.set_option("filter_i8", 0)
    for e in (e8, e16, e1):
        e.append(0, X)
    e8.set_kernel_timing(True)
    ref = e1.search_host(Q, k, metric=metric)
    for B in (1, 5):
        for eng, want_kernel in ((e8, 3), (e16, 2)):
            s, g, c = eng.search_host(Q[:B], k, metric=metric)
            if eng is e8:
                assert eng.stats()["last_kernel"] == want_kernel
            np.testing.assert_array_equal(g, ref[1][:B])
            np.testing.assert_array_equal(s.view(np.uint32), ref[0][:B].view(np.uint32))
            np.testing.assert_array_equal(c, ref[2][:B])
    for e in (e8, e16, e1):
        e.close()


@pytest.mark.parametrize("dim", [5, 20, 100, 130, 250, 1000, 1030, 1536, 3072, 8200])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_small_batch_kernels_on_ragged_dims(built_lib, dim, dtype):
    """The small-batch filter kernel (B <= 16) on dims that are not a multiple of 16 / 128: the last K block of the
    int8 (128 dims per block) and of the bf16 (64 dims per block) operand is partial, the shadow rows are padded
    (ld8 / ld16) and TMA zero-fills past the logical width.  Both shadows must stay bit-identical to the scan, for
    every metric, single queries and a 7-query batch, k = 10 and k = 100.  Wide rows (3072: the large text-embedding
    models; 8200: the in-kernel re-score stages the fp32 queries in several groups) ride along."""
    import wdbx_b200

    rng = np.random.default_rng(1000 + dim)
    n = 30011 if dim <= 1100 else 9001
    X = rng.standard_normal((n, dim), dtype=np.float32)
    X[17] = 0.0                                         # a zero row
    Q = rng.standard_normal((7, dim), dtype=np.float32)
    engines = {}
    for name, env in (("i8", {"WDBX_B200_GEMM_MIN_BATCH": "1"}),
                      ("bf16", {"WDBX_B200_GEMM_MIN_BATCH": "1", "WDBX_B200_FILTER_I8": "0"}),
                      ("scan", {"WDBX_B200_GEMM_MIN_BATCH": "0"})):
        os.environ.update(env)
        try:
            engines[name] = wdbx_b200.Engine(device=0, dim=dim, dtype=dtype, num_segments=1)
        finally:
            for key in env:
                os.environ.pop(key, None)
        engines[name].append(0, X[:5000])
        engines[name].append(0, X[5000:])              # a second append: the shadows grow behind the rows
        engines[name].set_kernel_timing(True)
    Xs = oracle.bf16_round(X) if dtype == "bf16" else X
    for metric in ("cosine", "ip", "l2"):
        for queries, k in ((Q[:1], 10), (Q, 10), (Q[2:5], 100)):
            want = engines["scan"].search_host(queries, k, metric=metric)
            assert engines["scan"].stats()["last_kernel"] == 1
            for name, kernel in (("i8", 3), ("bf16", 2)):
                got = engines[name].search_host(queries, k, metric=metric)
                if not (dtype == "bf16" and name == "bf16"):   # a bf16 store without the int8 shadow scans its own rows
                    assert engines[name].stats()["last_kernel"] == kernel, (name, metric, engines[name].stats()["last_kernel"])
                np.testing.assert_array_equal(got[2], want[2])
                np.testing.assert_array_equal(got[1], want[1])
                np.testing.assert_array_equal(got[0].view(np.uint32), want[0].view(np.uint32))
            if dim <= 1100 or metric == "cosine":       # (the fp64 oracle over wide rows is the slow part)
                _oracle_check(Xs, queries[:2], k, metric, *(w[:2] for w in want))
    for e in engines.values():
        e.close()


@pytest.mark.parametrize("B", [7, 40])   # small-batch kernel (int8 shadow) / 128-query kernel (bf16 shadow)
def test_non_finite_and_extreme_values_match_the_scan(built_lib, B):
    """+-inf / NaN elements, rows whose squared norm overflows fp32, denormal rows, and queries of the same kinds:
    "the caller's problem, but must not crash"  -- and whatever the scan returns for them, the filter
    paths must return the same bits (such rows are always candidates: their bound is inf / NaN)."""
    rng = np.random.default_rng(4242)
    n, dim, k = 20000, 96, 12
    X = rng.standard_normal((n, dim), dtype=np.float32)
    X[10, 3] = np.inf
    X[11, 4] = -np.inf
    X[12] = 1e38                      # |x|^2 overflows
    X[13] = 1e-40                     # denormals: |x|^2 underflows to 0
    X[14, 7] = np.nan
    X[15, 1], X[15, 2] = np.inf, -np.inf
    X[16] = -3e38
    X[17, 0] = 3.3e38
    X[5000:5010] *= 1e18              # large but finite rows
    X[6000:6010] *= 1e-18
    Q = rng.standard_normal((B, dim), dtype=np.float32)
    Q[1] *= 1e30
    Q[2] *= 1e-30
    Q[3, 5] = np.nan
    Q[4, 6] = np.inf
    Q[5] = 1e-40
    Q[6] = 3e38
    e2, e1 = _engine(dim, gemm_min_batch=1), _engine(dim, gemm_min_batch=0)
    for e in (e2, e1):
        e.append(0, X)
        e.set_kernel_timing(True)
    with np.errstate(all="ignore"):
        for metric in ("cosine", "ip", "l2"):
            s2, g2, c2 = e2.search_host(Q, k, metric=metric)
            assert e2.stats()["last_kernel"] in (2, 3)
            s1, g1, c1 = e1.search_host(Q, k, metric=metric)
            assert e1.stats()["last_kernel"] == 1
            np.testing.assert_array_equal(c2, c1, err_msg=metric)
            for b in range(B):
                # rows tied at -inf / NaN-as--inf may come in any order past the last finite score; everything before must agree
                fin = np.isfinite(s1[b]) | (s1[b] == np.inf)
                np.testing.assert_array_equal(g2[b][fin], g1[b][fin], err_msg=f"{metric} query {b}")
                np.testing.assert_array_equal(s2[b].view(np.uint32), s1[b].view(np.uint32), err_msg=f"{metric} query {b}")
    e2.close(); e1.close()

