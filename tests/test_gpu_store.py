"""GPU: the public API (WDBX / VectorStore / VectorIndex facades) over the real engine reproduces
the reference's golden outputs and the reference's own test assertions (tests/test_core.py)."""
import asyncio
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import wdbx_b200  # noqa: E402
from oracle import exact_search as oracle  # noqa: E402
from tests import golden_checks as gc  # noqa: E402


def make_store(dim, shards, **cfg):
    cfg.setdefault("GPU_STRICT", True)
    return wdbx_b200.VectorStore(dim, tempfile.mkdtemp(), num_shards=shards, config=wdbx_b200.WDBXConfig(cfg))


def test_ramp_golden(built_lib, golden):
    for case in golden["ramp"]:
        gc.check_ramp(make_store, case)


def test_self_query_golden(built_lib, golden):
    for case in golden["self_query"]:
        gc.check_self_query(make_store, case)


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_random_golden(built_lib, golden, idx):
    gc.check_random(make_store, golden["random"][idx])


def test_quickstart_c1_through_public_api(built_lib):
    """BASELINE.json configs[0]: WDBX(vector_dimension=384, num_shards=2), 10k vectors, limit=5, cosine."""
    rng = np.random.default_rng(1234)
    X = rng.standard_normal((10000, 384), dtype=np.float32)
    Q = rng.standard_normal((6, 384), dtype=np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        db = wdbx_b200.WDBX(vector_dimension=384, num_shards=2, data_dir=tmp, enable_gpu=True, log_level="WARNING")

        async def go():
            await db.initialize()
            ids = [f"doc{i}" for i in range(X.shape[0])]
            n = db.vector_store.batch_store({ids[i]: X[i] for i in range(5000)}, {i_: {"n": j} for j, i_ in enumerate(ids)})
            assert n == 5000
            for i in range(5000, 5050):
                await db.vector_store_async(X[i].tolist(), {"n": i}, id=ids[i])
            db.vector_store.batch_store({ids[i]: X[i] for i in range(5050, 10000)})
            assert db.count_vectors() == 10000
            ost = oracle.OracleStore(384, 1)
            for i in range(X.shape[0]):
                ost.add(0, ids[i], X[i])
            for b in range(Q.shape[0]):
                got = db.vector_search(Q[b].tolist(), limit=5)
                want = ost.search(Q[b], 5)
                gc.assert_same_results([(g[0], g[1], None) for g in got], [(w[0], w[1], None) for w in want])
                got_a = await db.vector_search_async(Q[b].tolist(), limit=5)
                assert got_a == got
            br = db.vector_search_batch(Q, limit=5)
            assert br.ids() == [[r[0] for r in db.vector_search(Q[b].tolist(), limit=5)] for b in range(Q.shape[0])]
            # self query / dimension error / stats (tests/test_core.py:135-142, :245-247, :341)
            r = db.vector_search(X[77].tolist(), limit=1)
            assert r[0][0] == "doc77" and r[0][1] > 0.99
            with pytest.raises(ValueError, match="dimension mismatch"):
                db.vector_search([0.0] * 3)
            st = db.get_stats()
            assert len(st["indices"]) == 2 and st["gpu"]["engine"]["kernel_launches"] > 0
            vec, meta = db.get_vector("doc5001")
            np.testing.assert_array_equal(np.asarray(vec, np.float32), X[5001])
            await db.shutdown()

        asyncio.run(go())


def test_crud_on_device(built_lib):
    st = make_store(4, 2)
    assert st.store("a", [1, 0, 0, 0], {"t": 1}) and st.store("b", [0, 1, 0, 0]) and st.store("c", [1, 1, 0, 0])
    assert st.store("a", [0, 0, 1, 0], {"t": 2}) and st.count() == 3
    assert st.get("a") == ([0.0, 0.0, 1.0, 0.0], {"t": 2})
    assert [r[0] for r in st.search([0, 0, 1, 0], limit=1)] == ["a"]
    assert st.delete("c") and len(st.search([1, 1, 0, 0], limit=10)) == 2
    assert st.clear() == 2 and st.search([0, 1, 0, 0]) == []
    assert st.store("a", [1, 0, 0, 0]) and st.search([1, 0, 0, 0], limit=5)[0][:2] == ("a", 1.0)
    st.close()


def test_metrics_and_bf16_through_store(built_lib):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((3000, 96), dtype=np.float32)
    Q = rng.standard_normal((3, 96), dtype=np.float32)
    for metric, dtype in (("ip", "fp32"), ("l2", "fp32"), ("ip", "bf16")):
        st = make_store(96, 2, GPU_METRIC=metric, GPU_DTYPE=dtype)
        st.bulk_load(X)
        Xs = oracle.bf16_round(X) if dtype == "bf16" else X
        res = st.search_batch(Q, limit=10)
        for b in range(3):
            rep = oracle.check_topk(Xs, Q[b], metric, 10, res.gids[b], res.scores[b])
            assert rep["hard_mismatch"] == 0 and rep["recall"] == 1.0 and rep["max_err_over_tol"] <= 1.0, rep
        st.close()


def test_engine_errors_follow_reference_convention(built_lib):
    st = make_store(4, 1, GPU_STRICT=False)
    st.store("a", [1, 0, 0, 0])
    st.engine.close()  # break the engine: reference convention = log + [] (indexing.py:1028-1030)
    assert st.search([1, 0, 0, 0]) == []
    assert st.store("b", [0, 1, 0, 0]) is False


def test_async_micro_batching_on_device(built_lib):
    st = make_store(96, 2, GPU_BATCH_MAX=8, GPU_BATCH_WINDOW_US=5000)
    rng = np.random.default_rng(4)
    X = rng.standard_normal((20000, 96), dtype=np.float32)
    st.bulk_load(X)
    Q = rng.standard_normal((40, 96), dtype=np.float32)
    want = [st.search(Q[b].tolist(), limit=10) for b in range(40)]
    n0 = st.engine.stats()["searches"]

    async def go():
        return await asyncio.gather(*[st.search_async(Q[b].tolist(), limit=10) for b in range(40)])

    got = asyncio.run(go())
    assert got == want                                   # multi-query passes are bit-identical to single
    used = st.engine.stats()["searches"] - n0
    assert used == st._batcher.batches and used <= 12    # 40 requests -> a handful of launches
    st.close()


def test_persistence_roundtrip_on_device(built_lib):
    d = tempfile.mkdtemp()
    rng = np.random.default_rng(6)
    X = rng.standard_normal((5000, 64), dtype=np.float32)
    for dtype in ("fp32", "bf16"):
        dd = d + dtype
        cfg = wdbx_b200.WDBXConfig({"GPU_STRICT": True, "GPU_DTYPE": dtype})
        st = wdbx_b200.VectorStore(64, dd, num_shards=3, config=cfg)
        st.bulk_load(X[:3000], id_prefix="b")
        st.batch_store({f"e{i}": X[3000 + i] for i in range(2000)}, {f"e{i}": {"i": i} for i in range(2000)})
        st.delete("e7"); st.delete("b3")
        Q = rng.standard_normal((4, 64), dtype=np.float32)
        want = [st.search(Q[b].tolist(), limit=20) for b in range(4)]
        asyncio.run(st.shutdown())
        st2 = wdbx_b200.VectorStore(64, dd, num_shards=3, config=cfg)
        assert st2.count() == 4998
        assert [st2.search(Q[b].tolist(), limit=20) for b in range(4)] == want   # bit-identical after reload
        st2.close()


def test_opt_in_prefilter_on_device(built_lib):
    rng = np.random.default_rng(12)
    n, dim = 30000, 96
    X = rng.standard_normal((n, dim), dtype=np.float32)
    pre = make_store(dim, 3, GPU_PREFILTER=True)
    ref = make_store(dim, 3)
    meta = {f"r{i}": {"i": i, "tag": "a" if i % 97 == 0 else "b"} for i in range(n)}
    for st in (pre, ref):
        st.batch_store({f"r{i}": X[i] for i in range(n)}, meta)
    match = np.array([i % 97 == 0 for i in range(n)])
    for b in range(3):
        q = rng.standard_normal(dim).astype(np.float32)
        got = pre.search(q.tolist(), limit=10, filter_metadata={"tag": "a"})
        rows, sc = oracle.topk_desc(oracle.scores_fp64(X, q, "cosine"), 10, dead=~match)
        assert [g[0] for g in got] == [f"r{r}" for r in rows]          # full k among the matching rows
        np.testing.assert_allclose([g[1] for g in got], sc, rtol=1e-5, atol=1e-6)
        assert len(ref.search(q.tolist(), limit=10, filter_metadata={"tag": "a"})) < 10   # reference semantics truncate
        hi = pre.search(q.tolist(), limit=10, threshold=float(sc[4] + sc[5]) / 2, filter_metadata={"tag": "a"})
        assert [g[0] for g in hi] == [f"r{r}" for r in rows[:5]]       # threshold pushed into the kernel
    pre.close(); ref.close()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("route", ["filter", "scan"])
def test_concurrent_searches_appends_and_deletes(built_lib, route, monkeypatch):
    """The reference runs index searches on a 4-thread pool next to `add` with no locks of its own (indexing.py:692,
    :1045-1048; SURVEY.md 8b "the engine must serialise append-vs-search itself").  Reader threads search while a
    writer stores and deletes: every answer must be a well-formed top-10 of SOME state in between (sorted, distinct,
    known ids, right metadata), the segments grow in place under the searches (the shadows with them on the filter
    route), and once the writer is done every reader's query returns exactly the oracle's list for the final store."""
    import threading

    monkeypatch.setenv("WDBX_B200_GEMM_MIN_BATCH", "1" if route == "filter" else "0")
    rng = np.random.default_rng(99)
    dim, n0, n_new = 128, 40000, 240
    X = rng.standard_normal((n0 + n_new, dim), dtype=np.float32)
    Q = rng.standard_normal((3, dim), dtype=np.float32)
    X[n0 + 5] = Q[0] * 2.0                  # a new row that takes over first place for reader 0
    st = make_store(dim, 2)
    st.bulk_load(X[:n0], id_prefix="v")
    errors, done = [], threading.Event()

    def reader(t):
        try:
            ql, last = Q[t].tolist(), None
            first = st.search(ql, limit=10)
            while True:
                fin = done.is_set()
                res = st.search(ql, limit=10)
                ids = [r[0] for r in res]
                assert len(res) == 10 and len(set(ids)) == 10, ids
                assert all(a[1] >= b[1] for a, b in zip(res, res[1:])), res
                for vid, _, meta in res:
                    assert vid[0] in "vw" and vid[1:].isdigit(), vid
                    assert meta == ({"i": int(vid[1:])} if vid[0] == "w" else {}), (vid, meta)
                last = res
                if fin:
                    break
            out[t] = (first, last)
        except BaseException as e:   # noqa: BLE001
            errors.append(repr(e))

    def writer():
        try:
            for i in range(n_new):
                assert st.store(f"w{i}", X[n0 + i].tolist(), {"i": i})
                if i % 6 == 0:
                    assert st.delete(f"v{i}")
        except BaseException as e:   # noqa: BLE001
            errors.append(repr(e))
        finally:
            done.set()

    out = {}
    threads = [threading.Thread(target=reader, args=(t,)) for t in range(3)] + [threading.Thread(target=writer)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(150)
    assert not errors, errors[:3]
    assert not any(th.is_alive() for th in threads)
    dead = np.zeros(n0 + n_new, bool)
    dead[np.arange(0, n_new, 6)] = True
    names = [f"v{i}" for i in range(n0)] + [f"w{i}" for i in range(n_new)]
    for t in range(3):
        rows, _ = oracle.topk_desc(oracle.scores_fp64(X, Q[t], "cosine"), 10, dead=dead)
        assert [r[0] for r in out[t][1]] == [names[r] for r in rows], t
        assert st.search(Q[t].tolist(), limit=10) == out[t][1]
    assert out[0][1][0][0] == "w5"
    assert st.count() == n0 + n_new - len(np.arange(0, n_new, 6))
    st.close()
