"""GPU, >= 2 devices: the SPMD store (one process per GPU) over NCCL and over the fused NVLink
exchange returns exactly what a single GPU returns.  Skipped on one-GPU boxes."""
import json
import os
import socket
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    rng = np.random.default_rng(5)
    X = rng.standard_normal((40000, 384), dtype=np.float32)
    X[123] = X[77]
    Q = rng.standard_normal((6, 384), dtype=np.float32)
    return X, Q


def _worker(rank, world, port, outdir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import wdbx_b200

    torch.cuda.set_device(rank)
    ctx = wdbx_b200.DistContext.from_env(rank)
    X, Q = _data()
    out = {}
    # route: "filter" = bf16-shadow filter + refine + stand-alone NVLink exchange kernel (what large stores use),
    # "scan" = K1 with the exchange fused into its last CTA; both must agree with one GPU
    for route, fused in (("filter", True), ("filter", False), ("scan", True), ("scan", False)):
        os.environ["WDBX_B200_SHADOW_MIN_MB"] = "0" if route == "filter" else "-1"
        st = wdbx_b200.VectorStore(384, tempfile.mkdtemp(), num_shards=2, dist=ctx,
                                   config=wdbx_b200.WDBXConfig({"GPU_STRICT": True, "GPU_FUSED_EXCHANGE": fused}))
        assert st._fused == fused
        st.bulk_load(X)
        st.delete("v5")
        res = [[(i, s) for i, s, _ in st.search(Q[b].tolist(), limit=10)] for b in range(Q.shape[0])]
        for _ in range(50):  # back-to-back collectives reuse the two exchange slots
            st.search(Q[0].tolist(), limit=10)
        qd = st.engine.upload(Q)
        dev = st.search_device(qd, 10)
        torch.cuda.synchronize()
        out[route + ("_fused" if fused else "_nccl")] = {"res": res, "dev_gids": dev["gids"].cpu().tolist(),
                                              "filtered": [(i, s) for i, s, _ in st.search(Q[1].tolist(), limit=5, filter_metadata={"x": 1})],
                                              "tie": [i for i, _, _ in st.search(X[77].tolist(), limit=2)]}
        st.close()
    os.environ.pop("WDBX_B200_SHADOW_MIN_MB", None)
    Path(outdir, f"rank{rank}.json").write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_match_one(built_lib):
    import torch
    import torch.multiprocessing as mp
    import wdbx_b200

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    X, Q = _data()
    one = wdbx_b200.VectorStore(384, tempfile.mkdtemp(), num_shards=2, dist=wdbx_b200.DistContext(0, 1, 0),
                                config=wdbx_b200.WDBXConfig({"GPU_STRICT": True}))
    one.bulk_load(X)
    one.delete("v5")
    want = json.loads(json.dumps([[(i, s) for i, s, _ in one.search(Q[b].tolist(), limit=10)] for b in range(Q.shape[0])]))
    one.close()
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(2, _free_port(), outdir), nprocs=2, join=True)
        got = [json.loads(Path(outdir, f"rank{r}.json").read_text()) for r in range(2)]
    for r in range(2):
        for mode in ("filter_fused", "filter_nccl", "scan_fused", "scan_nccl"):
            assert got[r][mode]["res"] == want, (r, mode)
            assert got[r][mode]["tie"] == ["v77", "v123"]
            assert got[r][mode]["filtered"] == []
    assert got[0] == got[1]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("route", ["filter", "scan"])
def test_single_process_group_matches_one_gpu(built_lib, route, monkeypatch):
    """GPU_DEVICES: ONE ordinary process drives 2 GPUs (no torchrun): ids and score BITS equal the one-GPU
    store's on every route -- on-device NVLink exchange (B <= 8), peer-copy + merge kernel (B = 40, k = 200),
    per-shard lists (metadata post-filter), opt-in pre-filter, the async micro-batcher, device-resident search."""
    import asyncio

    import torch
    import wdbx_b200
    from oracle import exact_search as oracle

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("WDBX_B200_SHADOW_MIN_MB", "0" if route == "filter" else "-1")
    X, Q = _data()
    Q40 = np.random.default_rng(9).standard_normal((40, 384), dtype=np.float32)
    meta = {f"v{i}": {"even": i % 2 == 0} for i in range(0, 4000)}

    def build(**cfg):
        st = wdbx_b200.VectorStore(384, tempfile.mkdtemp(), num_shards=2, dist=wdbx_b200.DistContext(0, 1, 0),
                                   config=wdbx_b200.WDBXConfig(dict(GPU_STRICT=True, **cfg)))
        st.bulk_load(X)
        st.metadata.update(meta)
        st.delete("v5")
        st.store("late", (X[9] * 2).tolist(), {"even": True})
        return st

    def probe(st):
        out = {"single": [[(i, np.float32(s).view(np.uint32).item()) for i, s, _ in st.search(Q[b].tolist(), limit=10)]
                          for b in range(Q.shape[0])]}
        for name, queries, k in (("b6", Q, 10), ("b40", Q40, 10), ("k200", Q[:3], 200)):
            r = st.search_batch(queries, k)
            out[name] = (r.gids.tolist(), r.scores.view(np.uint32).tolist(), r.counts.tolist())
        out["post"] = st.search(Q[1].tolist(), limit=7, filter_metadata={"even": True})
        out["tie"] = [i for i, _, _ in st.search(X[77].tolist(), limit=2)]
        out["thr"] = st.search(Q[2].tolist(), limit=10, threshold=0.12)
        out["get"] = st.get("late")[0]
        return out

    one = build()
    want = probe(one)
    multi = build(GPU_DEVICES="0,1")
    rows = multi.engine.stats()["rows_per_device"]
    assert sum(rows) == 40001 and abs(rows[0] - rows[1]) <= 2      # striped over both devices
    got = probe(multi)
    assert got == want
    assert want["tie"] == ["v77", "v123"]
    # oracle check of the one-GPU answer itself (so "equal" means "right")
    Xall = np.concatenate([X, X[9:10] * 2])
    ids = [f"v{i}" for i in range(X.shape[0])] + ["late"]
    for b in range(Q.shape[0]):
        sc = oracle.scores_fp32(Xall, Q[b], "cosine")
        sc[5] = -np.inf                                              # deleted
        top, _ = oracle.topk_desc(sc, 10)
        assert [i for i, _ in want["single"][b]] == [ids[r] for r in top]

    async def burst(st):
        return await asyncio.gather(*[st.search_async(Q[b % 6].tolist(), limit=10) for b in range(24)])
    assert asyncio.run(burst(multi)) == asyncio.run(burst(one))
    assert multi._batcher.batches < 24
    # device-resident search: queries and results on the first device, no host round trip
    qd = multi.engine.upload(Q)
    dev = multi.search_device(qd, 10)
    ref = one.search_device(one.engine.upload(Q), 10)
    torch.cuda.synchronize()
    assert torch.equal(dev["gids"].cpu(), ref["gids"].cpu()) and torch.equal(dev["keys"].cpu(), ref["keys"].cpu())
    # opt-in pre-filter (bitmaps split per device) + threshold push-down
    pre_m, pre_1 = build(GPU_DEVICES=[0, 1], GPU_PREFILTER=True), build(GPU_PREFILTER=True)
    f = {"even": True}
    a, b_ = pre_m.search(Q[3].tolist(), limit=9, threshold=0.05, filter_metadata=f), pre_1.search(Q[3].tolist(), limit=9, threshold=0.05, filter_metadata=f)
    assert a == b_ and 0 < len(a) <= 9 and all(m.get("even") for _, _, m in a)
    for st in (one, multi, pre_m, pre_1):
        st.close()


@pytest.mark.timeout(300)
def test_exchange_batches_larger_than_the_scan_block(built_lib):
    """at small dimensions the streaming scan scores fewer queries per block than the fused exchange
    admits (dim 64 fp32: one), so a batch of 2..8 queries must run as consecutive collective passes instead of
    failing with ERR_LIMIT.  Single-process group over 2 GPUs, scan route and filter route."""
    import torch
    import wdbx_b200

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(12)
    for dim in (64, 128):
        X = rng.standard_normal((30000, dim), dtype=np.float32)
        Q = rng.standard_normal((7, dim), dtype=np.float32)
        for shadow in ("-1", "0"):
            os.environ["WDBX_B200_SHADOW_MIN_MB"] = shadow
            try:
                one = wdbx_b200.VectorStore(dim, tempfile.mkdtemp(), dist=wdbx_b200.DistContext(0, 1, 0),
                                            config=wdbx_b200.WDBXConfig({"GPU_STRICT": True}))
                two = wdbx_b200.VectorStore(dim, tempfile.mkdtemp(), dist=wdbx_b200.DistContext(0, 1, 0),
                                            config=wdbx_b200.WDBXConfig({"GPU_STRICT": True, "GPU_DEVICES": "0,1"}))
            finally:
                os.environ.pop("WDBX_B200_SHADOW_MIN_MB", None)
            for st in (one, two):
                st.bulk_load(X)
            a, b = one.search_batch(Q, 10), two.search_batch(Q, 10)          # B = 7 <= 8: the on-device exchange
            np.testing.assert_array_equal(a.gids, b.gids)
            np.testing.assert_array_equal(a.scores.view(np.uint32), b.scores.view(np.uint32))
            one.close(); two.close()
