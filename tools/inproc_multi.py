"""Single-process multi-GPU e2e (no torchrun): WDBX(enable_gpu=True, config={"GPU_DEVICES": ...}) over C3
(10M x 768 fp32 cosine, top-10, batch 1).  Times `vector_search` (host list in, tuples out) and a burst of
concurrent `vector_search_async` calls (micro-batcher), checks ids against a chunked fp64 checker, prints JSON.
Usage: python tools/inproc_multi.py [n_gpus] [rows].  Not the bench."""
import asyncio, json, os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
sys.path.insert(0, str(ROOT))
import torch
import wdbx_b200
import bench

G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_ROWS
DIM, K = bench.DIM, bench.K
db = wdbx_b200.WDBX(vector_dimension=DIM, num_shards=1, data_dir=tempfile.mkdtemp(), enable_gpu=True,
                    config={"GPU_DEVICES": list(range(G)) if G > 1 else None, "GPU_STRICT": True})
store = db.vector_store
dev0 = torch.device("cuda", 0)
store.engine.reserve(0, ROWS)
for c in range((ROWS + bench.CHUNK - 1) // bench.CHUNK):
    m = min(bench.CHUNK, ROWS - c * bench.CHUNK)
    store.bulk_load(bench._gen_chunk(dev0, c, m), id_prefix=f"c{c}_")     # striped over the G devices by MultiEngine
Qd = bench._gen_queries(dev0)
Qlists = [Qd[i].cpu().tolist() for i in range(bench.N_QUERIES)]
for i in range(5):
    db.vector_search(Qlists[i], limit=K)
n = 200
t0 = time.perf_counter()
for i in range(n):
    res = db.vector_search(Qlists[i % bench.N_QUERIES], limit=K)
sync_s = time.perf_counter() - t0
dev_ms = store.engine.stats()["last_search_ms"]


async def burst(m):
    return await asyncio.gather(*[db.vector_search_async(Qlists[i % bench.N_QUERIES], limit=K) for i in range(m)])

asyncio.run(burst(64))
t0 = time.perf_counter()
got = asyncio.run(burst(512))
async_s = time.perf_counter() - t0
# parity: fp64 checker (world = 1 view of the whole matrix) through the device-resident group search
chk = bench._Fp64Checker(Qd, 1, 0, dev0, ROWS)
par = chk.check(lambda q: store.search_device(q, K), range(16))
ids_sync = [[r[0] for r in db.vector_search(Qlists[i], limit=K)] for i in range(4)]
ids_async = [[r[0] for r in got[i]] for i in range(4)]
print(json.dumps({
    "config": f"C3 {ROWS} x {DIM} fp32 cosine top-{K}, ONE process, {G} GPU(s) (GPU_DEVICES), no torchrun",
    "rows_per_device": store.engine.stats().get("rows_per_device", [ROWS]),
    "vector_search_qps": n / sync_s, "vector_search_ms": sync_s / n * 1e3, "device_ms_last_search": dev_ms,
    "vector_search_async_burst512_qps": 512 / async_s,
    "batcher": {"batches": store._batcher.batches, "requests": store._batcher.requests},
    "parity": par, "async_equals_sync": ids_sync == ids_async}))
db.close()
