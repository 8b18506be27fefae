"""Scratch: throughput of K1 with B queries per call (device-resident), C3 shape by default."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, wdbx_b200, ctypes as C
from wdbx_b200 import _lib

n, dim = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10_000_000, 768)
eng = wdbx_b200.Engine(0, dim, "fp32", 1)
eng.reserve(0, n)
g = torch.Generator(device="cuda").manual_seed(1)
done = 0
while done < n:
    m = min(1 << 20, n - done); eng.append(0, torch.randn((m, dim), generator=g, device="cuda")); done += m
lib = _lib.load_library()
def tune(warps, stages, U, qpp):
    # queries_per_pass is only settable through the environment at create time; poke via set_tuning + env not possible,
    # so this probe recreates nothing: it relies on WDBX_B200_QUERIES_PER_PASS for non-default QB.
    eng.set_tuning(warps, stages, U, 0, -1)
for B in (1, 2, 4, 8, 16, 64):
    q = torch.randn((B, dim), device="cuda")
    for (w, st, U) in [(0, 0, 0), (8, 2, 2), (8, 1, 4), (8, 2, 1)]:
        tune(w, st, U, 0)
        out = eng.search(q, 10, "cosine")
        for _ in range(2): eng.search(q, 10, "cosine", out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 10 if B <= 8 else 3
        e0.record()
        for _ in range(it): eng.search(q, 10, "cosine", out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print(f"B={B:3d} warps={w} stages={st} U={U}: {ms:8.3f} ms/call  {B/ms*1e3:9.1f} QPS  ({n*dim*4*max(1,(B+7)//8)/ms/1e6:7.0f} GB/s streamed)", flush=True)
eng.close()
