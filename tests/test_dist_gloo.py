"""CPU, world_size 2 over gloo: the SPMD path of VectorStore (row striping, all-gather of packed
keys, k-way merge) returns exactly what a single rank returns."""
import json
import os
import socket
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dataset():
    rng = np.random.default_rng(42)
    X = rng.standard_normal((301, 16), dtype=np.float32)
    X[40] = X[7]  # tie across ranks
    Q = rng.standard_normal((3, 16), dtype=np.float32)
    return X, Q


def _run_store(store):
    X, Q = _dataset()
    ids = [f"id{i}" for i in range(X.shape[0])]
    meta = {ids[i]: {"i": i, "even": i % 2 == 0} for i in range(len(ids))}
    store.batch_store({ids[i]: X[i] for i in range(200)}, meta)
    for i in range(200, X.shape[0]):
        store.store(ids[i], X[i].tolist(), meta[ids[i]])
    store.delete("id5")
    store.store("id9", X[11].tolist(), {"i": 9, "even": False})  # overwrite
    out = {"plain": [], "filtered": [], "batch": None, "get": None}
    for b in range(Q.shape[0]):
        out["plain"].append([(i, s) for i, s, _ in store.search(Q[b].tolist(), limit=10)])
        out["filtered"].append([(i, s) for i, s, _ in store.search(Q[b].tolist(), limit=6, filter_metadata={"even": True})])
    out["tie"] = [(i, s) for i, s, _ in store.search(X[7].tolist(), limit=3)]
    out["batch"] = store.search_batch(Q, limit=10).as_lists()
    out["get"] = store.get("id40")[0]
    out["shard0"] = store.indices[0].search(Q[0], limit=5)
    out["count"] = store.count()
    return out


def _worker(rank, world, port, outdir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import wdbx_b200
    from tests.fake_engine import FakeEngine

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = wdbx_b200.DistContext(rank, world, rank)
    store = wdbx_b200.VectorStore(16, tempfile.mkdtemp(), num_shards=3, dist=ctx, _engine_factory=FakeEngine)
    out = _run_store(store)
    out["local_rows"] = store.engine.stats()["rows_total"]
    Path(outdir, f"rank{rank}.json").write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_matches_single_rank():
    import wdbx_b200
    from tests.fake_engine import FakeEngine

    single = wdbx_b200.VectorStore(16, tempfile.mkdtemp(), num_shards=3, dist=wdbx_b200.DistContext(0, 1, 0),
                                   _engine_factory=FakeEngine)
    want = json.loads(json.dumps(_run_store(single)))
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(2, _free_port(), outdir), nprocs=2, join=True)
        got = [json.loads(Path(outdir, f"rank{r}.json").read_text()) for r in range(2)]
    rows = [g.pop("local_rows") for g in got]
    assert sum(rows) == 301 and abs(rows[0] - rows[1]) <= 3  # striped: balanced within one row per shard
    assert got[0] == got[1] == want
    assert [i for i, _ in want["tie"][:2]] == ["id7", "id40"]


# ---------------------------------------------------------------------------------------------- persistence
def _probe(store):
    X, Q = _dataset()
    return {"plain": [[(i, s) for i, s, _ in store.search(Q[b].tolist(), limit=10)] for b in range(Q.shape[0])],
            "tie": [(i, s) for i, s, _ in store.search(X[7].tolist(), limit=3)],
            "get": store.get("id40")[0], "count": store.count()}


def _restripe_worker(rank, world, port, data_dir, outdir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import wdbx_b200
    from tests.fake_engine import FakeEngine

    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = wdbx_b200.WDBXConfig({"GPU_STRICT": True})
    store = wdbx_b200.VectorStore(16, data_dir, num_shards=3, dist=wdbx_b200.DistContext(rank, world, rank), config=cfg,
                                  _engine_factory=FakeEngine)     # saved by ONE rank: re-striped over two
    out = {"loaded": _probe(store), "local_rows": store.engine.stats()["rows_total"]}
    X, _ = _dataset()
    store.store("late", (X[3] * 2.0).tolist(), {"late": True})
    store.delete("id12")
    out["after"] = _probe(store)
    assert store.save()
    Path(outdir, f"rank{rank}.json").write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_saved_store_is_restriped_across_world_sizes():
    """A store saved by W ranks loads on W' ranks (rows re-striped from the saved partitions) and back."""
    import wdbx_b200
    from tests.fake_engine import FakeEngine

    cfg = wdbx_b200.WDBXConfig({"GPU_STRICT": True})
    one = lambda d: wdbx_b200.VectorStore(16, d, num_shards=3, dist=wdbx_b200.DistContext(0, 1, 0), config=cfg,
                                          _engine_factory=FakeEngine)
    with tempfile.TemporaryDirectory() as data_dir, tempfile.TemporaryDirectory() as outdir:
        st = one(data_dir)
        _run_store(st)
        want_loaded = json.loads(json.dumps(_probe(st)))
        assert st.save()
        st.close()
        mp.spawn(_restripe_worker, args=(2, _free_port(), data_dir, outdir), nprocs=2, join=True)
        got = [json.loads(Path(outdir, f"rank{r}.json").read_text()) for r in range(2)]
        rows = [g.pop("local_rows") for g in got]
        assert sum(rows) == 301 and abs(rows[0] - rows[1]) <= 3
        assert got[0] == got[1] and got[0]["loaded"] == want_loaded
        assert not list(Path(data_dir).glob("shard_*/rows.rank*of1.npy"))      # superseded partitions removed
        back = one(data_dir)                                                      # two ranks -> one rank
        assert json.loads(json.dumps(_probe(back))) == got[0]["after"]
        assert back.count() == want_loaded["count"]                              # +late, -id12
        # "late" = 2 * id3: same cosine, so the tie rule (lower insertion id first) orders them
        assert [r[0] for r in back.search(_dataset()[0][3].tolist(), limit=2)] == ["id3", "late"]
        back.close()


# ---------------------------------------------------------------------------------------------- random op sequences
def _random_ops(make_store, seed, steps=70):
    """A seeded random walk over the store's mutations (the rules of tests/test_store_model.py, without hypothesis so
    that every rank replays exactly the same sequence); after every step a probe of what the store answers."""
    rng = np.random.default_rng(seed)
    dim, S = 6, 3
    store = make_store()
    explicit = [f"e{i}" for i in range(12)]
    live, prefixes, transcript = set(), 0, []
    Q = rng.standard_normal((2, dim)).astype(np.float32)

    def vec():
        return rng.standard_normal(dim).astype(np.float32)

    for step in range(steps):
        op = int(rng.integers(0, 10))
        if op == 0:
            vid = explicit[int(rng.integers(0, len(explicit)))]
            store.store(vid, vec().tolist(), {"g": int(rng.integers(0, 3))})
            live.add(vid)
        elif op == 1:
            ids = [explicit[i] for i in rng.choice(len(explicit), size=int(rng.integers(1, 5)), replace=False)]
            store.batch_store({v: vec().tolist() for v in ids}, {v: {"g": 1} for v in ids[::2]})
            live.update(ids)
        elif op == 2:
            n = int(rng.integers(1, 12))
            store.bulk_load(np.stack([vec() for _ in range(n)]), id_prefix=f"p{prefixes}_")
            live.update(f"p{prefixes}_{i}" for i in range(n))
            prefixes += 1
        elif op == 3 and live:
            vid = sorted(live)[int(rng.integers(0, len(live)))]
            transcript.append(("delete", store.delete(vid)))
            live.discard(vid)
        elif op == 4 and live:
            vid = sorted(live)[int(rng.integers(0, len(live)))]
            store.update_metadata(vid, {"g": 2, "u": step})
        elif op == 5:
            s = int(rng.integers(0, S))
            store.indices[s].clear()
            live = {v for v in live if store._locate(v) is not None}
        elif op == 6:
            vid = explicit[int(rng.integers(0, len(explicit)))]
            store.indices[int(rng.integers(0, S))].add(vid, vec())
            live.add(vid)
        elif op == 7 and live:
            vid = sorted(live)[int(rng.integers(0, len(live)))]
            removed = store.indices[int(rng.integers(0, S))].remove(vid)
            transcript.append(("index_remove", removed))
            if removed:
                live.discard(vid)
        elif op == 8 and step % 3 == 0:
            assert store.save()
            store.close()
            store = make_store()
        elif op == 9 and step % 7 == 0:
            store.clear()
            live, prefixes = set(), prefixes   # (prefix numbers keep growing: never reused here)
        # (scores to 4 decimals: the numpy double sums a stripe and the whole matrix in different orders; the device
        # engine's scores are bit-identical for every rank count, tests/test_gpu_multi.py)
        probe = {"count": store.count(), "live": len(live)}
        for b in range(2):
            probe[f"q{b}"] = [(i, round(s, 4), m) for i, s, m in store.search(Q[b].tolist(), limit=5)]
            probe[f"f{b}"] = [(i, round(s, 4)) for i, s, _ in store.search(Q[b].tolist(), limit=4, filter_metadata={"g": 1})]
        if live:
            vid = sorted(live)[0]
            got = store.get(vid)
            probe["get"] = [vid, None if got is None else [round(x, 6) for x in got[0]], None if got is None else got[1]]
        transcript.append((step, op, probe))
    assert store.count() == len(live)
    store.close()
    return transcript


def _ops_worker(rank, world, port, data_dir, outdir, seed):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import wdbx_b200
    from tests.fake_engine import FakeEngine

    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = wdbx_b200.WDBXConfig({"GPU_STRICT": True})
    make = lambda: wdbx_b200.VectorStore(6, data_dir, num_shards=3, dist=wdbx_b200.DistContext(rank, world, rank), config=cfg,
                                         _engine_factory=FakeEngine)
    out = _random_ops(make, seed)
    Path(outdir, f"rank{rank}.json").write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,seed", [(2, 11), (3, 12)])
def test_random_op_sequences_match_a_single_rank(world, seed):
    """The same seeded walk of mutations (explicit / batch / bulk rows, deletes, shard clears, index-level add / remove,
    save + reload, clear) on W gloo ranks and on one rank: every probe after every step is identical -- the striping of
    rows, overwrites on the owning rank, tombstones and the persisted partitions never leak into an answer."""
    import wdbx_b200
    from tests.fake_engine import FakeEngine

    cfg = wdbx_b200.WDBXConfig({"GPU_STRICT": True})
    with tempfile.TemporaryDirectory() as d1:
        want = json.loads(json.dumps(_random_ops(
            lambda: wdbx_b200.VectorStore(6, d1, num_shards=3, dist=wdbx_b200.DistContext(0, 1, 0), config=cfg,
                                          _engine_factory=FakeEngine), seed)))
    with tempfile.TemporaryDirectory() as data_dir, tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_ops_worker, args=(world, _free_port(), data_dir, outdir, seed), nprocs=world, join=True)
        got = [json.loads(Path(outdir, f"rank{r}.json").read_text()) for r in range(world)]
    for r in range(world):
        assert len(got[r]) == len(want)
        for a, b in zip(got[r], want):
            assert a == b, (r, a, b)
