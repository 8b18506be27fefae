"""Where does the host path spend its time?  search_host wall time vs device time (events)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import numpy as np, torch
import wdbx_b200

for n, dim in [(10_000, 384), (1_000_000, 384), (10_000_000, 768)]:
    eng = wdbx_b200.Engine(0, dim, "fp32", 1)
    g = torch.Generator(device="cuda").manual_seed(1)
    done = 0
    while done < n:
        m = min(1 << 20, n - done)
        eng.append(0, torch.randn((m, dim), generator=g, device="cuda"))
        done += m
    Q = np.random.default_rng(0).standard_normal((64, dim)).astype(np.float32)
    for i in range(5): eng.search_host(Q[i], 10)
    walls, devs = [], []
    for i in range(50):
        t0 = time.perf_counter(); eng.search_host(Q[i % 64], 10); walls.append(time.perf_counter() - t0)
        devs.append(eng.stats()["last_search_ms"])
    qd = torch.from_numpy(Q).cuda()
    out = eng.search(qd[0:1], 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): eng.search(qd[i % 64: i % 64 + 1], 10, out=out)
    e1.record(); torch.cuda.synchronize()
    # one-at-a-time with sync (latency, not throughput)
    lat = []
    for i in range(50):
        t0 = time.perf_counter(); eng.search(qd[i % 64: i % 64 + 1], 10, out=out); torch.cuda.synchronize(); lat.append(time.perf_counter() - t0)
    print(f"{n}x{dim}: search_host wall median {np.median(walls)*1e3:.3f} ms (min {min(walls)*1e3:.3f}), device part {np.median(devs):.3f} ms; "
          f"back-to-back device search {e0.elapsed_time(e1)/50:.3f} ms; launch+sync latency {np.median(lat)*1e3:.3f} ms", flush=True)
    eng.close()
