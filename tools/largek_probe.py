"""k = 1000 on C2 (1M x 384 fp32 cosine, single query): the visualisation caller's shape
(wdbx/utils/visualization.py:493-498).  Device-timed whole search; prints JSON.  Not the bench."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch
import wdbx_b200

res = []
for n, dim, ks in ((1_000_000, 384, (10, 100, 1000)), (10_000_000, 768, (100, 1000))):
    eng = wdbx_b200.Engine(0, dim, "fp32", 1)
    eng.reserve(0, n)
    g = torch.Generator(device="cuda").manual_seed(1)
    done = 0
    while done < n:
        m = min(1 << 20, n - done)
        eng.append(0, torch.randn((m, dim), generator=g, device="cuda"))
        done += m
    qs = torch.randn((8, 1, dim), device="cuda")
    for k in ks:
        out = eng.search(qs[0], k, "cosine")
        for i in range(5):
            eng.search(qs[i % 8], k, "cosine", out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            eng.search(qs[i % 8], k, "cosine", out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        r = {"rows": n, "dim": dim, "k": k, "ms_per_query": ms, "qps": 1e3 / ms,
             "stored_row_gbs": n * dim * 4 / ms / 1e6}
        res.append(r)
        print(json.dumps(r), flush=True)
    eng.close()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "largek_probe.json").write_text(json.dumps(res, indent=1))
