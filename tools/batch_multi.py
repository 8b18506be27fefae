"""Batched configs under torchrun (rows striped over ranks): C3 B=1024 (10M x 768 cosine) and C5 (5M x 1536 l2,
B=4096).  Every rank runs K2b on its rows, NCCL all-gathers the [B,k] keys, K3 merges.  Rank 0 prints JSON and
checks a sample of queries against a single-rank fp64 checker assembled with all_gather."""
import json, os, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, torch.distributed as dist
import wdbx_b200

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
ctx = wdbx_b200.DistContext.from_env(local)
dev = torch.device("cuda", local)
out = []
for name, total, dim, metric, B in (("C3 10M x 768 cosine B=1024", 10_000_000, 768, "cosine", 1024),
                                    ("C5 5M x 1536 l2 B=4096", 5_000_000, 1536, "l2", 4096)):
    st = wdbx_b200.VectorStore(dim, tempfile.mkdtemp(), num_shards=1, dist=ctx,
                               config=wdbx_b200.WDBXConfig({"GPU_METRIC": metric, "GPU_STRICT": True}))
    CH = 1_000_000
    for c in range(total // CH):
        g = torch.Generator(device=dev).manual_seed(500 + c)
        x = torch.randn((CH, dim), generator=g, device=dev)
        st.bulk_load({"local": x[rank::world].contiguous() if world > 1 else x, "total": CH}, id_prefix=f"c{c}_")
        del x
    Q = torch.randn((B, dim), generator=torch.Generator(device=dev).manual_seed(9), device=dev)
    res = st.search_device(Q, 10)
    res = st.search_device(Q, 10)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 5
    e0.record()
    for _ in range(it): res = st.search_device(Q, 10)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / it], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # checker: fp64 scores of 4 queries over this rank's rows, global top-10 through all_gather
    sel = [0, 1, B // 2, B - 1]
    Qd = Q[sel].double()
    best_s, best_g = [], []
    for c in range(total // CH):
        g = torch.Generator(device=dev).manual_seed(500 + c)
        x = torch.randn((CH, dim), generator=g, device=dev)
        xs = (x[rank::world] if world > 1 else x).double()
        if metric == "cosine":
            s = (Qd @ xs.T) / (Qd.norm(dim=1, keepdim=True) * xs.norm(dim=1)[None, :])
        else:
            s = -((Qd * Qd).sum(1, keepdim=True) - 2.0 * (Qd @ xs.T) + (xs * xs).sum(1)[None, :])
        v, i = torch.topk(s, 10, dim=1)
        best_s.append(v); best_g.append(c * CH + rank + i * world)
        del x, xs, s
    v, gsel = torch.cat(best_s, 1), torch.cat(best_g, 1)
    if world > 1:
        vs = [torch.empty_like(v) for _ in range(world)]; gs = [torch.empty_like(gsel) for _ in range(world)]
        dist.all_gather(vs, v); dist.all_gather(gs, gsel)
        v, gsel = torch.cat(vs, 1), torch.cat(gs, 1)
    top = torch.topk(v, 10, dim=1)
    want = torch.gather(gsel, 1, top.indices)
    ok = bool((want == res["gids"][sel]).all())
    if rank == 0:
        t = float(ms.item())
        out.append({"config": name, "n_gpus": world, "ms_per_batch": t, "qps": B / t * 1e3,
                    "useful_tflops_total": 2.0 * total * dim * B / t / 1e9, "ids_match_fp64_checker": ok})
    st.close()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
