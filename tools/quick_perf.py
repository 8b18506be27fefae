"""Scratch perf probe for K1 (device-resident queries, CUDA events).  Not the bench."""
import sys, time, itertools, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import numpy as np, torch
import wdbx_b200

def fill(eng, n, dim, seed=0, chunk=1 << 20):
    g = torch.Generator(device="cuda").manual_seed(seed)
    done = 0
    while done < n:
        m = min(chunk, n - done)
        x = torch.randn((m, dim), generator=g, device="cuda", dtype=torch.float32)
        eng.append(0, x)
        done += m

def timeit(eng, q, k, metric, iters=20, warm=3):
    out = eng.search(q, k, metric)
    for _ in range(warm): eng.search(q, k, metric, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): eng.search(q, k, metric, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    cfgs = [(1_000_000, 384, "fp32"), (10_000_000, 768, "fp32"), (12_500_000, 384, "bf16")]
    if len(sys.argv) > 1: cfgs = [cfgs[int(a)] for a in sys.argv[1:]]
    for n, dim, dt in cfgs:
        import os
        K = int(os.environ.get("QP_K", 100 if dt == "bf16" else 10))
        eng = wdbx_b200.Engine(0, dim, dt, 1)
        t0 = time.time(); fill(eng, n, dim); torch.cuda.synchronize()
        print(f"== {n}x{dim} {dt}: filled in {time.time()-t0:.1f}s", flush=True)
        q = torch.randn((1, dim), device="cuda")
        eb = 2 if dt == "bf16" else 4
        for metric in (("ip",) if dt == "bf16" else ("cosine",)):
            bytes_ = n * dim * eb + (4 * n if metric == "cosine" else 0)
            for warps, stages, U in [(0, 0, 0), (16, 2, 4), (12, 2, 4)]:
                try:
                    eng.set_tuning(warps, stages, U, 0, -1)
                except Exception as ex:
                    print("  skip", warps, stages, U, ex); continue
                ms = timeit(eng, q, K, metric)
                print(f"  {metric:6s} warps={warps:2d} stages={stages} U={U}: {ms*1e3:8.1f} us  {bytes_/ms/1e6:8.1f} GB/s  ({bytes_/ms/1e6/6533.5:.3f} of measured)", flush=True)
        eng.close()

if __name__ == "__main__":
    main()
