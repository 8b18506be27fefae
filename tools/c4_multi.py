"""C4 under torchrun: 100M x 384 bf16 inner-product top-100, rows striped over the ranks (12.5M per GPU at N=8).
Measures device-resident QPS with the fused NVLink exchange and with the NCCL path; rank 0 prints JSON."""
import json, os, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, torch.distributed as dist
import wdbx_b200

TOTAL = int(os.environ.get("C4_ROWS", 100_000_000)); DIM = 384; K = 100
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
ctx = wdbx_b200.DistContext.from_env(local)
dev = torch.device("cuda", local)
res = {}
for fused in (True, False):
    st = wdbx_b200.VectorStore(DIM, tempfile.mkdtemp(), num_shards=1, dist=ctx, config=wdbx_b200.WDBXConfig(
        {"GPU_DTYPE": "bf16", "GPU_METRIC": "ip", "GPU_STRICT": True, "GPU_FUSED_EXCHANGE": fused}))
    per = TOTAL // world
    st.engine.reserve(0, per)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    done, c = 0, 0
    while done < per:
        m = min(1 << 20, per - done)
        x = torch.randn((m, DIM), generator=g, device=dev)
        st.bulk_load({"local": x, "total": m * world}, id_prefix=f"c{c}_")
        done += m; c += 1
    Q = torch.randn((16, DIM), generator=torch.Generator(device=dev).manual_seed(7), device=dev)
    qs = [Q[i:i + 1] for i in range(16)]
    for i in range(5): out = st.search_device(qs[i], K)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 200
    e0.record()
    for i in range(steps): out = st.search_device(qs[i % 16], K)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res["fused" if fused else "nccl"] = {"ms_per_query": float(ms.item()), "qps": 1e3 / float(ms.item()),
                                        "gids_q0": st.search_device(qs[0], K)["gids"][0, :5].tolist()}
    rows_local = st.engine.stats()["rows_total"]
    st.close()
if rank == 0:
    per_gpu_bytes = rows_local * DIM * 2
    for v in res.values():
        v["hbm_gbs_per_gpu"] = per_gpu_bytes / v["ms_per_query"] / 1e6
    assert res["fused"]["gids_q0"] == res["nccl"]["gids_q0"]
    print(json.dumps({"config": f"C4 {TOTAL} x {DIM} bf16 ip top-{K}, {world} GPU(s), rows striped", "rows_per_gpu": rows_local, **res}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
