"""C4 under torchrun: 100M x 384 bf16 inner-product top-100, rows striped over the ranks (12.5M per GPU at N=8).
Measures device-resident QPS with the fused NVLink exchange and with the NCCL path, and checks BOTH routes against a
chunked fp64 checker over the bf16-rounded rows of all ranks (4 queries, global top-100); rank 0 prints JSON."""
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
    # ---- fp64 checker (not the product path): this rank's rows are re-generated from the same generator stream
    NQ = 4
    Qc = Q[:NQ].double()
    gchk = torch.Generator(device=dev).manual_seed(100 + rank)
    best_s, best_g, done, base = [], [], 0, 0
    while done < per:
        m = min(1 << 20, per - done)
        x = torch.randn((m, DIM), generator=gchk, device=dev).to(torch.bfloat16).double()
        s = Qc @ x.T                                             # [NQ, m] inner products of the stored (bf16) rows
        v, i = torch.topk(s, min(K, m), dim=1)
        best_s.append(v)
        best_g.append(base + rank + world * i)                   # bulk_load: position n = rank + world * i of the chunk
        done += m
        base += m * world
        del x, s
    v, gsel = torch.cat(best_s, 1), torch.cat(best_g, 1)
    if world > 1:
        vs = [torch.empty_like(v) for _ in range(world)]
        gs = [torch.empty_like(gsel) for _ in range(world)]
        dist.all_gather(vs, v)
        dist.all_gather(gs, gsel)
        v, gsel = torch.cat(vs, 1), torch.cat(gs, 1)
    top = torch.topk(v, K, dim=1)
    want_g, want_s = torch.gather(gsel, 1, top.indices).cpu(), top.values.cpu()
    ok_ids, ok_sets, worst = True, True, 0.0
    for b in range(NQ):
        o = st.search_device(qs[b], K)
        gg, gsc = o["gids"][0].cpu(), o["scores"][0].cpu().double()
        ok_ids = ok_ids and bool((gg == want_g[b]).all())
        ok_sets = ok_sets and set(gg.tolist()) == set(want_g[b].tolist())
        worst = max(worst, float(((gsc - want_s[b]).abs() / want_s[b].abs().clamp_min(1.0)).max()))
    res["fused" if fused else "nccl"] = {"ms_per_query": float(ms.item()), "qps": 1e3 / float(ms.item()),
                                        "gids_q0": st.search_device(qs[0], K)["gids"][0, :5].tolist(),
                                        "parity": {"checked_queries": NQ, "ids_match_fp64_checker": ok_ids,
                                                   "id_sets_match": ok_sets, "max_rel_score_err": worst}}
    rows_local = st.engine.stats()["rows_total"]
    st.close()
if rank == 0:
    per_gpu_bytes = rows_local * DIM * 2
    for v in res.values():
        v["hbm_gbs_per_gpu"] = per_gpu_bytes / v["ms_per_query"] / 1e6
    assert res["fused"]["gids_q0"] == res["nccl"]["gids_q0"]
    assert all(v["parity"]["id_sets_match"] and v["parity"]["max_rel_score_err"] < 1e-5 for v in res.values()), res
    print(json.dumps({"config": f"C4 {TOTAL} x {DIM} bf16 ip top-{K}, {world} GPU(s), rows striped", "rows_per_gpu": rows_local, **res}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
