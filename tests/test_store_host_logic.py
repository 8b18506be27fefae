"""CPU: host-side logic of wdbx_b200.VectorStore / WDBX against the reference's golden outputs,
with the device engine replaced by the numpy double (tests/fake_engine.py)."""
import asyncio
import tempfile

import numpy as np
import pytest

import wdbx_b200
from tests import golden_checks as gc
from tests.fake_engine import FakeEngine, FakeGroup, pack_keys, unpack_keys


def make_store(dim, shards, **cfg):
    return wdbx_b200.VectorStore(dim, tempfile.mkdtemp(), num_shards=shards, config=wdbx_b200.WDBXConfig(cfg),
                                 dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)


def test_ramp_golden(golden):
    for case in golden["ramp"]:
        gc.check_ramp(make_store, case)


def test_self_query_golden(golden):
    for case in golden["self_query"]:
        gc.check_self_query(make_store, case)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_random_golden(golden, idx):
    gc.check_random(make_store, golden["random"][idx])


def test_crud_semantics():
    st = make_store(4, 2)
    assert st.store("a", [1, 0, 0, 0], {"t": 1}) and st.store("b", [0, 1, 0, 0]) and st.store("c", [1, 1, 0, 0])
    assert st.count() == 3
    assert st.get("a") == ([1.0, 0.0, 0.0, 0.0], {"t": 1}) and st.get("zz") is None
    # duplicate id overwrites in place (HNSW intent, indexing.py:370-375), no ghost row
    assert st.store("a", [0, 0, 1, 0], {"t": 2})
    assert st.count() == 3 and st.get("a")[0] == [0.0, 0.0, 1.0, 0.0] and st.get("a")[1] == {"t": 2}
    assert [r[0] for r in st.search([0, 0, 1, 0], limit=1)] == ["a"]
    # deleted ids are never returned (SURVEY.md 8c decision 1)
    assert st.delete("c") and not st.delete("c") and st.count() == 2
    assert "c" not in [r[0] for r in st.search([1, 1, 0, 0], limit=10)]
    assert len(st.search([1, 1, 0, 0], limit=10)) == 2
    assert st.update_metadata("b", {"x": 1}) and not st.update_metadata("nope", {})
    assert st.search([0, 1, 0, 0], limit=1)[0] == ("b", 1.0, {"x": 1})
    # limit <= 0 and empty store
    assert st.search([0, 1, 0, 0], limit=0) == []
    assert st.clear() == 2 and st.count() == 0 and st.search([0, 1, 0, 0]) == []
    # ids keep working after clear
    assert st.store("a", [1, 0, 0, 0]) and st.search([1, 0, 0, 0], limit=5)[0][0] == "a"
    stats = st.get_stats()
    assert stats["use_gpu"] is True and stats["num_shards"] == 2 and len(stats["indices"]) == 2
    assert sum(i["size"] for i in stats["indices"]) == 1


def test_batch_and_bulk_apis():
    st = make_store(8, 3)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((100, 8), dtype=np.float32)
    assert st.bulk_load(X, id_prefix="v") == 100
    more = {f"m{i}": rng.standard_normal(8).astype(np.float32) for i in range(10)}
    assert st.batch_store(more) == 10 and st.count() == 110
    allX = np.concatenate([X, np.stack(list(more.values()))])
    ids = [f"v{i}" for i in range(100)] + list(more)
    Q = rng.standard_normal((4, 8), dtype=np.float32)
    res = st.search_batch(Q, limit=7)
    from oracle import exact_search as oracle
    for b in range(4):
        rows, sc = oracle.topk_desc(oracle.scores_fp32(allX, Q[b], "cosine"), 7)
        assert res.ids()[b] == [ids[r] for r in rows]
    assert st.get("v17")[0] == pytest.approx(X[17].tolist())
    assert st.delete("v17") and st.get("v17") is None and "v17" not in sum(st.search_batch(Q, 110).ids(), [])
    with pytest.raises(ValueError):
        st.bulk_load(X, id_prefix="v")


def test_facade_contract():
    with tempfile.TemporaryDirectory() as tmp:
        import wdbx_b200.wdbx as facade

        class _W(facade.WDBX):
            def _init_vector_store(self):
                self._store = wdbx_b200.VectorStore(self.vector_dim, self.data_dir, self.num_shards, config=self.config,
                                                    dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)

        db = _W(vector_dimension=4, num_shards=2, data_dir=tmp, config={"WDBX_TEST_OPTION": "x"}, enable_plugins=False)
        assert db.config.get("WDBX_TEST_OPTION") == "x"

        async def go():
            await db.initialize()
            vid = await db.vector_store_async([0.1, 0.2, 0.3, 0.4], {"source": "test"})
            res = await db.vector_search_async([0.1, 0.2, 0.3, 0.4], limit=1)
            assert res[0][0] == vid and res[0][1] > 0.99 and res[0][2] == {"source": "test"}   # test_core.py:167-174
            assert db.vector_search([0.1, 0.2, 0.3, 0.4], limit=1)[0][0] == vid             # test_core.py:135-142
            with pytest.raises(ValueError, match="dimension mismatch"):                      # test_core.py:245-247
                db.vector_search([0.1, 0.2, 0.3])
            with pytest.raises(ValueError, match="dimension mismatch"):
                await db.vector_store_async([0.1], {})
            assert db.get_vector("nonexistent_id") is None
            assert db.count_vectors() == 1 and db.delete_vector(vid) and db.count_vectors() == 0
            st = db.get_stats()
            assert st["gpu_enabled"] is True and len(st["indices"]) == 2 and st["total_vectors"] == 0
            await db.shutdown()

        asyncio.run(go())


def test_no_cpu_fallback():
    with tempfile.TemporaryDirectory() as tmp:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            wdbx_b200.VectorStore(4, tmp, use_gpu=False)


def test_key_packing_roundtrip():
    s = np.array([1.5, -2.0, 0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30], np.float32)
    g = np.arange(8)
    k = pack_keys(s, g)
    sc, gi = unpack_keys(k)
    assert list(gi) == list(g)
    np.testing.assert_array_equal(sc[:6], np.array([1.5, -2.0, 0.0, 0.0, np.inf, -np.inf], np.float32))
    assert np.isneginf(sc[6])
    order = np.argsort(-k.astype(np.float64), kind="stable")
    assert list(order[:3]) == [4, 0, 7] and k[2] > k[3]  # +0 == -0 -> lower gid first


def test_async_micro_batching_matches_sync():
    st = make_store(8, 2, GPU_BATCH_MAX=8, GPU_BATCH_WINDOW_US=20000)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((300, 8), dtype=np.float32)
    st.batch_store({f"r{i}": X[i] for i in range(300)}, {f"r{i}": {"i": i} for i in range(300)})
    Q = rng.standard_normal((20, 8), dtype=np.float32)
    want = [st.search(Q[b].tolist(), limit=3 + b % 4, threshold=0.2 if b % 5 == 0 else 0.0) for b in range(20)]
    launches0 = st.engine.launches

    async def go():
        return await asyncio.gather(*[st.search_async(Q[b].tolist(), limit=3 + b % 4,
                                                      threshold=0.2 if b % 5 == 0 else 0.0) for b in range(20)])

    got = asyncio.run(go())
    assert got == want
    assert st._batcher.requests == 20 and st._batcher.batches <= 6      # 20 requests coalesced into few launches
    assert st.engine.launches - launches0 == st._batcher.batches
    # filtered requests bypass the batcher and keep the reference semantics
    f = asyncio.run(st.search_async(Q[0].tolist(), limit=5, filter_metadata={"i": {"$lt": 100}}))
    assert f == st.search(Q[0].tolist(), limit=5, filter_metadata={"i": {"$lt": 100}})
    # a query of the wrong dimension: logged and [] like the reference's index-level convention (indexing.py:1028-1030);
    # GPU_STRICT raises (the WDBX facade validates dimensions itself and always raises, wdbx.py:323-326)
    assert asyncio.run(st.search_async([0.0] * 3)) == [] and st.search([0.0] * 3) == []
    st.strict = True
    with pytest.raises(ValueError, match="dimension mismatch"):
        asyncio.run(st.search_async([0.0] * 3))
    st.close()


def test_persistence_roundtrip():
    import tempfile as tf

    d = tf.mkdtemp()
    rng = np.random.default_rng(8)
    X = rng.standard_normal((120, 8), dtype=np.float32)

    def mk():
        return wdbx_b200.VectorStore(8, d, num_shards=3, config=wdbx_b200.WDBXConfig({"GPU_STRICT": True}),
                                     dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)

    st = mk()
    st.bulk_load(X[:60], id_prefix="b")
    st.batch_store({f"e{i}": X[60 + i] for i in range(60)}, {f"e{i}": {"i": i} for i in range(60)})
    st.delete("e7"); st.delete("b3")
    st.store("e9", X[0].tolist(), {"i": 900})     # overwrite
    Q = rng.standard_normal((5, 8), dtype=np.float32)
    want = [st.search(Q[b].tolist(), limit=12) for b in range(5)]
    want_f = st.search(Q[0].tolist(), limit=4, filter_metadata={"i": {"$gte": 30}})
    asyncio.run(st.shutdown())                     # reference: shutdown persists (vector_store.py:202-217)
    st2 = mk()
    assert st2.count() == 118 and st2.get("e7") is None and st2.get("b3") is None
    assert [st2.search(Q[b].tolist(), limit=12) for b in range(5)] == want
    assert st2.search(Q[0].tolist(), limit=4, filter_metadata={"i": {"$gte": 30}}) == want_f
    assert st2.get("e9")[1] == {"i": 900} and st2.get("b10")[0] == pytest.approx(X[10].tolist())
    # the store keeps working after a reload: new ids, deletes, another save/load cycle
    st2.store("new", X[5].tolist(), {"i": -1})
    assert st2.search(X[5].tolist(), limit=2)[0][0] in ("b5", "new")
    assert st2.save()
    st3 = mk()
    assert st3.count() == 119 and st3.search(X[5].tolist(), limit=3) == st2.search(X[5].tolist(), limit=3)
    st2.close(); st3.close()


def test_opt_in_prefilter_returns_full_k_among_matches():
    rng = np.random.default_rng(12)
    X = rng.standard_normal((400, 8), dtype=np.float32)
    meta = {f"r{i}": {"i": i, "tag": "a" if i % 7 == 0 else "b"} for i in range(400)}
    ref = make_store(8, 3)                       # reference post-filter semantics
    pre = make_store(8, 3, GPU_PREFILTER=True)   # opt-in device-side pre-filter
    for st in (ref, pre):
        st.batch_store({f"r{i}": X[i] for i in range(400)}, meta)
    q = rng.standard_normal(8).astype(np.float32).tolist()
    flt = {"tag": "a", "i": {"$gte": 100}}
    got = pre.search(q, limit=5, filter_metadata=flt)
    from oracle import exact_search as oracle
    match = np.array([i % 7 == 0 and i >= 100 for i in range(400)])
    rows, sc = oracle.topk_desc(oracle.scores_fp32(X, np.asarray(q, np.float32), "cosine"), 5, dead=~match)
    assert [g[0] for g in got] == [f"r{r}" for r in rows] and len(got) == 5
    assert all(m["tag"] == "a" and m["i"] >= 100 for _, _, m in got)
    assert len(ref.search(q, limit=5, filter_metadata=flt)) <= len(got)   # the post-filter truncates
    # threshold push-down and cache invalidation on mutation
    hi = pre.search(q, limit=50, threshold=0.3, filter_metadata=flt)
    assert all(s >= 0.3 for _, s, _ in hi)
    pre.delete(got[0][0])
    assert got[0][0] not in [g[0] for g in pre.search(q, limit=5, filter_metadata=flt)]
    pre.update_metadata(got[1][0], {"i": 0, "tag": "b"})
    assert got[1][0] not in [g[0] for g in pre.search(q, limit=5, filter_metadata=flt)]


def test_data_feeders_csv_jsonl(tmp_path):
    from wdbx_b200 import data_utils as du

    assert du.parse_vector("[1, 2.5, -3]") == [1.0, 2.5, -3.0]
    assert du.parse_vector("1, 2,3") == [1.0, 2.0, 3.0] and du.parse_vector("1 2 3") == [1.0, 2.0, 3.0]
    assert du.parse_vector("[1. 2.]") == [1.0, 2.0] and du.parse_vector({"embedding": [1, 2]}) == [1.0, 2.0]
    for bad in ("not a vector", "array([1., 2.])"):     # numpy's repr (with commas) is rejected by the reference too
        with pytest.raises(ValueError):
            du.parse_vector(bad)
    csv_path = tmp_path / "v.csv"
    csv_path.write_text('id,vec,lang\na,"[1,0,0,0]",en\nb,"0 1 0 0",de\nbad,"oops",xx\nc,"0,0,1,0",en\n')
    vecs, meta = du.load_vectors_from_csv(str(csv_path), "vec", id_column="id", metadata_columns=["lang"])
    assert list(vecs) == ["a", "b", "c"] and meta["b"] == {"lang": "de"} and vecs["c"] == [0.0, 0.0, 1.0, 0.0]
    vecs_i, meta_i = du.load_vectors_from_csv(str(csv_path), 1, id_column=0, metadata_columns=[2])
    assert vecs_i == vecs and meta_i["a"] == {"col_2 ": "en"}
    jl = tmp_path / "v.jsonl"
    jl.write_text('{"id": "x", "emb": [0, 0, 0, 1], "t": 1}\n{"id": "y", "emb": "1 1 0 0", "t": 2}\n{"id": "z", "t": 3}\n')
    vj, mj = du.load_vectors_from_jsonl(str(jl), "emb", id_field="id")
    assert list(vj) == ["x", "y"] and mj["y"] == {"id": "y", "t": 2}
    st = make_store(4, 2)
    assert du.ingest_csv(st, str(csv_path), "vec", id_column="id", metadata_columns=["lang"]) == 3
    assert du.ingest_jsonl(st, str(jl), "emb", id_field="id") == 2
    assert st.count() == 5 and st.search([0, 0, 0, 1], limit=1)[0][0] == "x"
    assert [r[0] for r in st.search([1, 0, 0, 0], limit=5, filter_metadata={"lang": "en"})] == ["a", "c"]
    st.close()


def test_list_query_conversion_matches_numpy():
    """The public API takes lists of Python floats (reference: vector_store.py:301); the fast struct-based
    conversion must produce exactly numpy's fp32 rounding and fall back to numpy for everything else."""
    st = wdbx_b200.VectorStore(8, tempfile.mkdtemp(), num_shards=1, dist=wdbx_b200.DistContext(0, 1, 0),
                               _engine_factory=FakeEngine)
    rng = np.random.default_rng(3)
    cases = [rng.standard_normal(8).tolist(), [0.1] * 8, [1, 2, 3, 4, 5, 6, 7, 8], [float("nan")] + [0.0] * 7,
             [1e-46] * 8, [np.float32(1.5)] * 8, [3.4028235677973366e38] * 8]
    for lst in cases:
        with np.errstate(over="ignore"):   # the last case rounds to inf (numpy warns, struct refuses -> numpy path)
            got, want = st._query_array(lst), np.array(lst, dtype=np.float32)
        assert got.dtype == np.float32 and got.shape == (8,)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), lst
    with np.errstate(over="ignore"):
        assert np.array_equal(st._query_array([1e39] * 8), np.array([1e39] * 8, dtype=np.float32))   # overflow -> inf via numpy
    assert st._query_array(np.arange(8)).dtype == np.float32                  # non-list input
    assert st._query_array([0.0] * 3).shape == (3,)                           # wrong length: numpy path, search() raises
    assert st.search([0.0] * 3) == []
    st.strict = True
    with pytest.raises(ValueError, match="dimension mismatch"):
        st.search([0.0] * 3)
    st.close()


def test_shard_clear_drops_bulk_ids_of_that_shard():
    """bulk ids of a cleared shard must not resolve to the (reused) row positions."""
    st = make_store(4, 2)
    X = np.eye(4, dtype=np.float32)[[0, 1, 2, 3, 0, 1]]
    assert st.bulk_load(X, id_prefix="v") == 6
    st.update_metadata("v0", {"m": 1})
    assert st.indices[0].clear()
    assert st.get("v0") is None and st.get("v2") is None and not st.delete("v0")
    assert st.get("v1")[0] == X[1].tolist()
    assert st.count() == 3 and "v0" not in st.metadata
    # a new id landing in shard 0 reuses row 0 of the shard: the stale bulk id must stay dead
    new = next(f"n{i}" for i in range(100) if st._get_shard_for_id(f"n{i}") == 0)
    assert st.store(new, [0, 0, 0, 1])
    assert st.get("v0") is None and not st.delete("v0") and st.get(new)[0] == [0.0, 0.0, 0.0, 1.0]
    assert st.count() == 4
    # the id can be stored again explicitly
    assert st.store("v0", [1, 0, 0, 0]) and st.get("v0")[0] == [1.0, 0.0, 0.0, 0.0]


def test_bulk_ids_are_canonical_and_prefix_collisions_rejected():
    """'v01' / unicode digits must not alias bulk row 1; explicit ids may not shadow bulk rows."""
    st = make_store(4, 1)
    X = np.eye(4, dtype=np.float32)
    st.bulk_load(X, id_prefix="v")
    assert st.get("v1") is not None
    assert st.get("v01") is None and not st.delete("v01") and st.get("v²") is None and st.get("v") is None
    assert st.store("v²", [1, 1, 0, 0]) and st.count() == 5
    st.store("w2", [0, 0, 1, 1])
    with pytest.raises(ValueError):
        st.bulk_load(X, id_prefix="w")
    assert st.bulk_load(X[:2], id_prefix="w") == 2      # w2 is outside the new range: no collision


def _multi_probe(st, X, Q):
    return {
        "plain": [[(i, s) for i, s, _ in st.search(Q[b].tolist(), limit=6)] for b in range(Q.shape[0])],
        "filtered": [(i, s) for i, s, _ in st.search(Q[0].tolist(), limit=4, filter_metadata={"even": True})],
        "batch": st.search_batch(Q, 5).as_lists(),
        "tie": [i for i, _, _ in st.search(X[7].tolist(), limit=2)],
        "get": st.get("v9")[0], "count": st.count(),
    }


def test_single_process_multi_device_store_matches_one_device(tmp_path):
    """GPU_DEVICES: one ordinary process, rows striped over G engines (multi_engine.py); every answer must equal
    the one-device store's, through CRUD, filters, the micro-batcher and persistence."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((157, 8), dtype=np.float32)
    X[40] = X[7]
    Q = rng.standard_normal((3, 8), dtype=np.float32)

    def build(path, **cfg):
        st = wdbx_b200.VectorStore(8, path, num_shards=3, config=wdbx_b200.WDBXConfig(dict(GPU_STRICT=True, **cfg)),
                                   dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)
        return st

    def fill(st):
        st.bulk_load(X[:100], id_prefix="v")
        st.batch_store({f"m{i}": X[100 + i] for i in range(57)}, {f"m{i}": {"even": i % 2 == 0} for i in range(57)})
        st.delete("v3")
        st.delete("m10")
        st.store("m2", (X[5] * 3).tolist(), {"even": True})   # overwrite in place on whichever device owns the row

    one = build(tmp_path / "one")
    fill(one)
    want = _multi_probe(one, X, Q)
    multi = build(tmp_path / "multi", GPU_DEVICES="0-2")
    assert multi.devices == [0, 1, 2] and multi._batcher is not None
    fill(multi)
    rows = multi.engine.stats()["rows_per_device"]
    assert sum(rows) == 157 and max(rows) - min(rows) <= 3        # striped: balanced within one row per shard
    assert _multi_probe(multi, X, Q) == want

    async def burst():
        return await asyncio.gather(*[multi.search_async(Q[b % 3].tolist(), limit=6) for b in range(12)])
    got = asyncio.run(burst())
    assert [[(i, s) for i, s, _ in r] for r in got] == [want["plain"][b % 3] for b in range(12)]
    assert multi._batcher.batches < 12                              # requests were coalesced
    # opt-in device pre-filter: the position bitmaps are split per device
    pre = build(tmp_path / "pre", GPU_DEVICES=[0, 1], GPU_PREFILTER=True)
    one_pre = build(tmp_path / "one_pre", GPU_PREFILTER=True)
    fill(pre)
    fill(one_pre)
    f = {"even": True}
    assert pre.search(Q[1].tolist(), limit=9, filter_metadata=f) == one_pre.search(Q[1].tolist(), limit=9, filter_metadata=f)
    assert len(pre.search(Q[1].tolist(), limit=9, filter_metadata=f)) == 9
    # persistence: saved by the 3-device store, loaded on one device and on two
    assert multi.save()
    multi.close()
    for cfg in ({}, {"GPU_DEVICES": "0,1"}):
        back = build(tmp_path / "multi", **cfg)
        assert _multi_probe(back, X, Q) == want
        back.close()


def test_multi_device_append_is_all_or_nothing():
    """A device that cannot make room fails the append BEFORE any stripe is written: the segment's positions stay
    aligned over the devices and the next append works (multi_engine.py, pre-flight reserve)."""
    from wdbx_b200.multi_engine import MultiEngine

    class Tight(FakeEngine):
        limit = None

        def reserve(self, segment, rows):
            self.reserved = max(getattr(self, "reserved", 0), rows)
            if self.limit is not None and rows > self.limit:
                raise MemoryError(f"device {self.device}: cannot hold {rows} rows")

    me = MultiEngine([0, 1, 2], 4, "fp32", 1, _engine_factory=Tight, _group_factory=FakeGroup)
    X = np.arange(40, dtype=np.float32).reshape(10, 4)
    assert me.append(0, X[:5]) == 0
    assert [e.rows[0].shape[0] for e in me.engines] == [2, 2, 1]
    assert [e.reserved for e in me.engines] == [2, 2, 1]
    me.engines[2].limit = 1                     # device 2 is full
    with pytest.raises(MemoryError):
        me.append(0, X[5:])                     # would put rows 5 and 8 on device 2
    assert [e.rows[0].shape[0] for e in me.engines] == [2, 2, 1] and me._seg_rows == [5]
    me.engines[2].limit = None
    assert me.append(0, X[5:]) == 5
    assert [e.rows[0].shape[0] for e in me.engines] == [4, 3, 3]
    np.testing.assert_array_equal(me.read_rows(0, 0, 10), X)
    me.close()


def test_gpu_devices_spec_parsing():
    from wdbx_b200.multi_engine import parse_devices

    assert parse_devices(None) is None and parse_devices("") is None
    assert parse_devices("0-3") == [0, 1, 2, 3] and parse_devices("0,2, 5") == [0, 2, 5]
    assert parse_devices([1, 0]) == [1, 0] and parse_devices(2) == [0, 1] and parse_devices("1-2,4") == [1, 2, 4]
    with pytest.raises(ValueError):
        parse_devices("0,0")


def test_searches_inside_an_append_resolve_ids_and_metadata(tmp_path):
    """Searches do not take the store lock (the reference runs them on a thread pool next to `add`,
    indexing.py:692).  A search that lands right after the device append -- before the host bookkeeping of that
    append is complete -- must already see the new rows under their ids, with their metadata (the reference sets
    metadata before index.add, vector_store.py:241-246), never under the str(gid) fallback."""
    seen = []

    class Hooked(FakeEngine):
        hook = None

        def append(self, segment, rows, gids=None):
            first = super().append(segment, rows, gids=gids)
            if Hooked.hook is not None:
                Hooked.hook()
            return first

    st = wdbx_b200.VectorStore(4, tmp_path, num_shards=2, config=wdbx_b200.WDBXConfig({"GPU_STRICT": True}),
                               dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=Hooked)
    q = [1.0, 0.0, 0.0, 0.0]

    def probe():
        located = set(st._loc)
        seen.append((st.search(q, limit=50), located))
        seen.append((st.search(q, limit=50, filter_metadata={"tag": "new"}), located))

    # later rows score higher, so a search inside their append returns THEM (k is clipped to the old live count)
    Hooked.hook = probe
    assert st.store("a", [1.0, 0.6, 0.0, 0.0], {"tag": "new"})
    assert st.batch_store({"b": [1.0, 0.5, 0.0, 0.0], "c": [1.0, 0.4, 0.0, 0.0]}, {"b": {"tag": "new"}, "c": {"tag": "new"}}) == 2
    assert st.bulk_load(np.asarray([[1.0, 0.3, 0.0, 0.0], [1.0, 0.2, 0.0, 0.0], [1.0, 0.1, 0.0, 0.0]], np.float32), id_prefix="k") == 3
    Hooked.hook = None
    known = {"a", "b", "c", "k0", "k1", "k2"}
    early = 0
    for res, located in seen:
        for vid, score, meta in res:
            assert vid in known, f"a search inside an append saw the fallback id {vid!r}"
            if vid in ("a", "b", "c"):
                assert meta == {"tag": "new"}
                early += vid not in located
            else:
                early += 1        # bulk rows are only ever probed inside their own bulk_load here
    assert early > 0, "no probe hit the window this test is about"
    # a failing append leaves nothing behind: no id, no metadata, and the next store works
    def boom():
        raise RuntimeError("device append failed")
    Hooked.hook = boom
    st.strict = False
    assert st.store("z", [0.0, 1.0, 0.0, 0.0], {"tag": "z"}) is False
    Hooked.hook = None
    assert "z" not in st.metadata and st._locate("z") is None
    st.close()


def test_filter_operator_ladder_matches_the_reference_grid(tmp_path):
    """tests/golden/filter_ops_golden.json holds what the reference's unmodified `VectorStore._matches_filter`
    (wdbx/core/vector_store.py:414-463) answers -- or raises -- on 660 (metadata, filter) pairs; ours must agree on
    every one, exceptions (by type) included."""
    import json
    from pathlib import Path

    g = json.loads((Path(__file__).resolve().parent / "golden" / "filter_ops_golden.json").read_text())
    st = wdbx_b200.VectorStore(4, tmp_path, num_shards=1, dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)

    def ours(f):
        try:
            return bool(st._matches_filter("id", f))
        except Exception as e:   # noqa: BLE001
            return "raises:" + type(e).__name__

    it = iter(g["outcomes"])
    n = 0
    for md in g["metadata"]:
        st.metadata = {"id": md}
        for f in g["filters"]:
            want = next(it)
            assert ours(f) == want, (md, f, want)
            n += 1
    st.metadata = {}
    for f, want in zip(g["filters"], g["no_entry"]):
        assert ours(f) == want, ("no metadata entry", f, want)
    assert n == len(g["outcomes"]) == 660
    st.close()


def test_feeders_match_the_reference_on_the_loader_grid(tmp_path):
    """tests/golden/loader_golden.json: the reference's unmodified parse_vector / load_vectors_from_csv /
    load_vectors_from_jsonl (wdbx/utils/data_utils.py) on 51 vector notations and 27 loader calls over small files
    (quoted multi-line cells, bad rows, duplicate ids, index vs name addressing, missing columns / fields / files);
    ours must return the same dictionaries and raise the same exception types."""
    import json
    import logging
    from pathlib import Path

    from wdbx_b200 import data_utils as du

    g = json.loads((Path(__file__).resolve().parent / "golden" / "loader_golden.json").read_text())

    def jsonable(x):
        if isinstance(x, dict):
            return {(k if isinstance(k, str) else f"<{k!r}>"): jsonable(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [jsonable(v) for v in x]
        if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
            return f"<{x!r}>"
        return x

    def outcome(fn, *a, **kw):
        try:
            return jsonable({"ok": fn(*a, **kw)})
        except Exception as e:   # noqa: BLE001
            return {"raises": type(e).__name__}

    # the parse inputs as they were before JSON flattened them (a tuple among them): the literal grid of the generator
    logging.disable(logging.CRITICAL)
    try:
        src = (Path(__file__).resolve().parent / "golden" / "make_loader_golden.py").read_text()
        ns = {}
        exec(src[src.index("PARSE = ["):src.index("CSV_FILES = {")], ns)   # the literal grid only, nothing from /root/reference
        assert jsonable(ns["PARSE"]) == g["parse_inputs"]
        for x, want in zip(ns["PARSE"], g["parse"]):
            assert outcome(du.parse_vector, x) == want, (x, want)
        for name, text in list(g["csv_files"].items()) + list(g["jsonl_files"].items()):
            with open(tmp_path / name, "w", encoding="utf-8", newline="") as f:
                f.write(text)
        for (name, kw), want in zip(g["csv_calls"], g["csv"]):
            assert outcome(du.load_vectors_from_csv, str(tmp_path / name), **kw) == want, (name, kw)
        for (name, kw), want in zip(g["jsonl_calls"], g["jsonl"]):
            assert outcome(du.load_vectors_from_jsonl, str(tmp_path / name), **kw) == want, (name, kw)
    finally:
        logging.disable(logging.NOTSET)


def test_operator_boundary_never_raises_like_the_reference(tmp_path):
    """Probed against the reference's unmodified VectorStore over the exact faiss stand-in (same call sequence, see
    DESIGN.md section 1): a query / vector of the wrong dimension at the VectorStore or VectorIndex level is logged and
    answered with [] / False (indexing.py:903-905, :1028-1030) -- only the WDBX facade raises (wdbx.py:323-326); limit <= 0
    gives []; a filter whose comparison is a type error raises TypeError out of search (no try around it there either);
    delete / update of unknown ids return False, get returns None.  Deliberate differences (DESIGN.md "decisions"): a
    rejected vector is not counted, a deleted row is never returned, a duplicate id overwrites."""
    st = wdbx_b200.VectorStore(4, tmp_path, num_shards=2, dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)
    assert st.search([1, 0, 0, 0], 5) == []
    assert st.store("a", [1, 0, 0, 0], {"t": 1}) is True and st.store("b", [0, 1, 0, 0]) is True
    assert st.store("w", [1, 2, 3]) is False and st.store("s", "abc") is False and st.store("e", []) is False
    assert st.count() == 2 and "w" not in st.metadata
    assert st.batch_store({"c": [1, 1, 0, 0], "d": [0, 0, 1, 0]}, {"c": {"t": 2}}) == 2
    assert st.batch_store({"x": [1, 1], "y": [0, 0, 1, 0], "z": "oops"}, {}) == 1          # only y is a usable vector
    assert st.count() == 5 and "x" not in st.metadata and st.get("x") is None
    ix = st.indices[0]
    assert ix.add("bad", np.zeros(3, np.float32)) is False
    assert ix.batch_add({"r1": np.zeros(4, np.float32), "r2": np.zeros(2, np.float32)}) is False   # ragged: never raises
    assert ix.search(np.zeros(3, np.float32), 3) == [] and ix.remove("nope") is False
    assert st.search([1, 0.1, 0, 0], 0) == [] and st.search([1, 0.1, 0, 0], -1) == []
    assert st.search([1, 0.1, 0], 3) == []
    assert [r[0] for r in st.search([1, 0, 0, 0], 10, 0.7)] == ["a", "c"]
    assert [r[0] for r in st.search([1, 0, 0, 0], 10, 0.0, {"t": {"$gte": 1}})] == ["a", "c"]
    with pytest.raises(TypeError):
        st.search([1, 0, 0, 0], 10, 0.0, {"t": {"$gt": "x"}})
    assert st.get("zz") is None and st.update_metadata("zz", {}) is False and st.delete("zz") is False
    assert st.delete("b") is True and st.delete("b") is False
    assert "b" not in [r[0] for r in st.search([0, 1, 0, 0], 10)]
    st.close()


def test_index_level_mutations_invalidate_the_prefilter_bitmaps(tmp_path):
    """Found by tests/test_store_model.py: a row added through the shard's index facade did not bump the store version,
    so a cached allow bitmap (GPU_PREFILTER) stayed in use although it was SHORTER than the segment -- on the device
    that is a read past the bitmap.  Every mutation path invalidates the cache now."""
    st = wdbx_b200.VectorStore(4, tmp_path, num_shards=2, config=wdbx_b200.WDBXConfig({"GPU_STRICT": True, "GPU_PREFILTER": True}),
                               dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)
    assert st.store("a", [1, 0, 0, 0], {"g": 1}) and st.store("b", [0.9, 0.1, 0, 0], {"g": 0})
    flt = {"g": 1}
    assert [r[0] for r in st.search([1, 0, 0, 0], limit=5, filter_metadata=flt)] == ["a"]      # bitmaps cached
    v0 = st._version
    assert st.indices[0].add("c", np.asarray([1, 0.01, 0, 0], np.float32)) and st._version > v0
    assert [r[0] for r in st.search([1, 0, 0, 0], limit=5, filter_metadata=flt)] == ["a"]      # c has no metadata: not allowed
    st.update_metadata("c", {"g": 1})
    assert [r[0] for r in st.search([1, 0, 0, 0], limit=5, filter_metadata=flt)] == ["a", "c"]
    assert st.indices[st._locate("c")[0]].remove("c") and "c" not in st.metadata               # no stale metadata either
    assert [r[0] for r in st.search([1, 0, 0, 0], limit=5, filter_metadata=flt)] == ["a"]
    st.close()


def test_batcher_mixed_limits_and_cancelled_requests(tmp_path):
    """One pass serves requests with different limits (a limit <= 0 gives [] as on the synchronous path -- it used to
    become a negative slice) and a request whose caller went away (cancelled future) does not disturb the others."""
    st = wdbx_b200.VectorStore(4, tmp_path, num_shards=2, config=wdbx_b200.WDBXConfig({"GPU_STRICT": True, "GPU_BATCH_WINDOW_US": 50000}),
                               dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)
    rng = np.random.default_rng(8)
    X = rng.standard_normal((40, 4), dtype=np.float32)
    st.bulk_load(X, id_prefix="v")
    q = rng.standard_normal(4).astype(np.float32)
    want = st.search(q.tolist(), limit=7)
    b = st._batcher
    n0 = b.batches
    futs = [b.submit(q, 7, 0.0), b.submit(q, -1, 0.0), b.submit(q, 0, 0.0), b.submit(q, 3, 0.0), b.submit(q, 100, want[2][1])]
    gone = b.submit(q, 5, 0.0)
    assert gone.cancel()
    tail = b.submit(-q, 2, 0.0)
    got = [f.result(timeout=10) for f in futs]
    assert got[0] == want and got[1] == [] and got[2] == [] and got[3] == want[:3] and got[4] == want[:3]
    assert [r[0] for r in tail.result(timeout=10)] == [r[0] for r in st.search((-q).tolist(), limit=2)]
    assert b.batches - n0 == 1                     # all of it was one pass
    st.close()


def test_async_front_end_completes_a_pass_with_one_loop_wakeup(tmp_path):
    """search_async hands the batcher an asyncio future; the futures of one pass are completed by ONE thread-safe
    call into their loop.  Cancelled waiters are skipped, engine errors follow the convention (strict: the awaiting
    coroutine sees the exception; otherwise logged and [])."""
    st = wdbx_b200.VectorStore(4, tmp_path, num_shards=1, config=wdbx_b200.WDBXConfig({"GPU_BATCH_WINDOW_US": 20000}),
                               dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)
    rng = np.random.default_rng(5)
    st.bulk_load(rng.standard_normal((30, 4), dtype=np.float32), id_prefix="v")
    qs = [rng.standard_normal(4).astype(np.float32).tolist() for _ in range(12)]
    want = [st.search(q, limit=4) for q in qs]

    async def burst():
        loop = asyncio.get_running_loop()
        calls = []
        orig = loop.call_soon_threadsafe
        loop.call_soon_threadsafe = lambda *a, **kw: (calls.append(a[0].__name__), orig(*a, **kw))[1]
        tasks = [asyncio.ensure_future(st.search_async(q, limit=4)) for q in qs]
        await asyncio.sleep(0)                 # everybody has submitted; the window (20 ms) is still open
        tasks[3].cancel()
        res = await asyncio.gather(*tasks, return_exceptions=True)
        loop.call_soon_threadsafe = orig
        return res, calls

    n0 = st._batcher.batches
    res, calls = asyncio.run(burst())
    assert isinstance(res[3], asyncio.CancelledError)
    assert [r for i, r in enumerate(res) if i != 3] == [w for i, w in enumerate(want) if i != 3]
    assert st._batcher.batches - n0 == 1 and calls.count("_complete_many") == 1
    # engine failure inside the pass
    st.engine.search_host = lambda *a, **kw: (_ for _ in ()).throw(RuntimeError("device lost"))
    assert asyncio.run(st.search_async(qs[0], limit=4)) == []
    st.strict = True
    with pytest.raises(RuntimeError, match="device lost"):
        asyncio.run(st.search_async(qs[0], limit=4))
    st.close()


@pytest.mark.parametrize("devices", [None, "0-2"])
def test_limits_above_the_engine_maximum_are_served_by_paging(tmp_path, devices, monkeypatch):
    """The reference returns up to `limit` rows (indexing.py:1005); the engine's widest list is MAX_K.  Larger limits
    are answered by successive exact passes over the rows not returned yet (masked with the pre-filter's allow bitmaps):
    the concatenation must be the exact ranking, duplicates (ties) and deleted rows included."""
    import wdbx_b200._lib as lib

    monkeypatch.setattr(lib, "MAX_K", 16)            # small pages: the walk below needs 1 + 87 // 16 of them
    cfg = {"GPU_STRICT": True}
    if devices:
        cfg["GPU_DEVICES"] = devices
    st = wdbx_b200.VectorStore(6, tmp_path, num_shards=3, config=wdbx_b200.WDBXConfig(cfg),
                               dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=FakeEngine)
    rng = np.random.default_rng(17)
    X = rng.standard_normal((90, 6), dtype=np.float32)
    if not devices:      # (the numpy double sums the stripes of a device group in different orders: its "ties" are not
        X[40] = X[7]     #  bit-exact across devices as the real engine's are, tests/test_gpu_multi.py)
        X[41] = X[7]     # a three-way tie somewhere in the ranking
    st.bulk_load(X[:60], id_prefix="v")
    st.batch_store({f"w{i}": X[60 + i] for i in range(30)}, {f"w{i}": {"i": i} for i in range(30)})
    for vid in ("v3", "w4", "v59"):
        st.delete(vid)
    q = rng.standard_normal(6).astype(np.float32)
    names = [f"v{i}" for i in range(60)] + [f"w{i}" for i in range(30)]
    dead = np.zeros(90, bool)
    dead[[3, 64, 59]] = True
    from oracle import exact_search as oracle
    rows, _ = oracle.topk_desc(oracle.scores_fp32(X, q, "cosine"), 90, dead=dead)
    want = [names[r] for r in rows]
    for limit in (17, 40, 87, 500):
        got = st.search(q.tolist(), limit=limit)
        assert [g[0] for g in got] == want[:limit], limit
        assert all(a[1] >= b[1] for a, b in zip(got, got[1:]))
        assert got[-1][2] == st.metadata.get(got[-1][0], {})
    assert [g[0] for g in st.search(q.tolist(), limit=16)] == want[:16]          # at the maximum: the ordinary path
    if not devices:
        assert [g[0] for g in st.search(X[7].tolist(), limit=30)][:3] == ["v7", "v40", "v41"]
    # threshold and filters keep their semantics (a filtered search stays one list of <= MAX_K per shard)
    thr = st.search(q.tolist(), limit=80, threshold=0.2)
    assert [g[0] for g in thr] == [g[0] for g in st.search(q.tolist(), limit=87) if g[1] >= 0.2]
    st.close()
