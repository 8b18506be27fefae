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
