"""Scratch: K2 (tcgen05) throughput. Usage: gemm_perf.py [n] [dim] [B...]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, wdbx_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
Bs = [int(a) for a in sys.argv[3:]] or [128, 1024]
os.environ["WDBX_B200_GEMM_MIN_BATCH"] = "1"
os.environ.setdefault("WDBX_B200_GEMM_MODE", "0")
eng = wdbx_b200.Engine(0, dim, "fp32", 1)
eng.reserve(0, n)
g = torch.Generator(device="cuda").manual_seed(1)
done = 0
while done < n:
    m = min(1 << 20, n - done); eng.append(0, torch.randn((m, dim), generator=g, device="cuda")); done += m
for metric in ("cosine",):
    for B in Bs:
        q = torch.randn((B, dim), device="cuda")
        out = eng.search(q, 10, metric)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 3
        e0.record()
        for _ in range(it): eng.search(q, 10, metric, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        fl = 2.0 * n * dim * B
        print(f"{metric} n={n} dim={dim} B={B}: {ms:9.3f} ms  {B/ms*1e3:10.1f} QPS  useful {fl/ms/1e9:8.1f} TFLOP/s  issued(3x) {3*fl/ms/1e9:8.1f} TFLOP/s", flush=True)
eng.close()
