"""CPU property tests (hypothesis): ranking-key order, top-k / merge algebra, shard striping.
These are the size-independent invariants the GPU kernels rely on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import exact_search as oracle
from tests.fake_engine import pack_keys, unpack_keys
from wdbx_b200.shard_map import ShardMap, shard_for_id

finite = st.floats(allow_nan=False, allow_infinity=False, width=32)
anyf = st.one_of(finite, st.sampled_from([float("inf"), float("-inf"), float("nan"), 0.0, -0.0]))


@settings(max_examples=200, deadline=None)
@given(st.lists(anyf, min_size=1, max_size=60))
def test_key_order_is_score_desc_then_gid_asc(scores):
    s = np.asarray(scores, np.float32)
    g = np.arange(len(s))
    keys = pack_keys(s, g)
    want_rows, _ = oracle.topk_desc(s, len(s))                            # score desc (NaN = -inf), row asc
    # float64 cannot hold 64-bit keys exactly: compare through a lexsort on the two halves instead
    hi, lo = (keys >> np.uint64(32)).astype(np.int64), (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    order = np.lexsort((-lo, -hi))
    assert list(order) == list(want_rows)
    sc, gi = unpack_keys(keys)
    assert list(gi) == list(g)
    ok = ~np.isnan(s)
    assert np.array_equal(sc[ok] + np.float32(0), s[ok] + np.float32(0))


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 40), st.integers(1, 6), st.integers(0, 2**31 - 1))
def test_topk_of_union_is_topk_of_partial_topks(k, parts, seed):
    """The identity behind per-warp -> per-CTA -> per-GPU -> global merging."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 300))
    s = rng.standard_normal(n).astype(np.float32)
    s[rng.integers(0, n, size=max(1, n // 10))] = s[0]                    # ties
    keys = pack_keys(s, np.arange(n))
    full = np.sort(keys)[::-1][:k]
    owner = rng.integers(0, parts, size=n)
    partial = np.concatenate([np.sort(keys[owner == p])[::-1][:k] for p in range(parts)])
    assert np.array_equal(np.sort(partial)[::-1][:k], full)


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 64), st.integers(1, 8), st.integers(0, 5000))
def test_striping_is_a_bijection_and_balanced(num_shards, world, total):
    m = ShardMap(num_shards, world)
    seen = set()
    for n in range(min(total, 400)):
        r, local = m.owner(n)
        assert 0 <= r < world and (r, local) not in seen
        seen.add((r, local))
    counts = [m.local_count(total, r) for r in range(world)]
    assert sum(counts) == total and max(counts) - min(counts) <= 1


@given(st.text(min_size=0, max_size=30), st.integers(1, 64))
def test_shard_for_id_is_deterministic_and_in_range(vid, shards):
    assert 0 <= shard_for_id(vid, shards) < shards and shard_for_id(vid, shards) == shard_for_id(vid, shards)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 50), st.integers(2, 24), st.integers(1, 12), st.sampled_from(["cosine", "ip", "l2"]),
       st.integers(0, 2**31 - 1))
def test_oracle_fp32_ranking_matches_fp64_outside_tie_window(n, dim, k, metric, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal(dim).astype(np.float32)
    rows, sc = oracle.topk_desc(oracle.scores_fp32(X, q, metric), k)
    rep = oracle.check_topk(X, q, metric, k, rows, sc)
    assert rep["hard_mismatch"] == 0 and rep["max_err_over_tol"] <= 1.0 and rep["sorted"]


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 90), st.integers(1, 8), st.integers(1, 8), st.integers(0, 2**31 - 1))
def test_restriping_saved_partitions_preserves_every_row(count, saved_world, world, seed):
    """Position n of a shard lives on rank n % W at local row n // W; loading W-rank partitions on W' ranks
    (VectorStore._load_partition) must hand every rank exactly its positions, in order."""
    import tempfile
    from pathlib import Path
    from types import SimpleNamespace

    from wdbx_b200.vector_store import VectorStore

    dim = 3
    full = np.random.default_rng(seed).standard_normal((count, dim)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "shard_0").mkdir()
        for r in range(saved_world):
            np.save(Path(d) / "shard_0" / f"rows.rank{r}of{saved_world}.npy", full[r::saved_world])
        me = SimpleNamespace(data_dir=Path(d), vector_dim=dim)
        got = [VectorStore._load_partition(me, 0, count, r, world, saved_world) for r in range(world)]
    for r in range(world):
        assert got[r].dtype == np.float32 and np.array_equal(got[r], full[r::world])
