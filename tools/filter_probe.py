"""Probe of the K2b path on ONE GPU: whole-step time, filter-kernel time (engine CUDA events) and
candidates handed to the exact refine, for a list of (rows, dim, metric, batch) cases.  rows = 1.25M x 768 is
the slice one rank holds at N = 8 of the 10M benchmark.  Writes gpurun_out/filter_probe.json.  Not the bench."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch
import wdbx_b200

PEAK = 6533.5
try:
    PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    pass


def fill(eng, n, dim, seed):
    eng.reserve(0, n)
    g = torch.Generator(device="cuda").manual_seed(seed)
    done = 0
    while done < n:
        m = min(1 << 20, n - done)
        eng.append(0, torch.randn((m, dim), generator=g, device="cuda"))
        done += m


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    cases = [(1_250_000, 768, "cosine", [1, 8]), (10_000_000, 768, "cosine", [1, 8, 64, 1024]),
             (1_000_000, 384, "cosine", [1]), (5_000_000, 1536, "l2", [1, 4096])]
    if len(sys.argv) > 1:
        cases = [cases[int(a)] for a in sys.argv[1:]]
    res = []
    for n, dim, metric, batches in cases:
        eng = wdbx_b200.Engine(0, dim, "fp32", 1)
        fill(eng, n, dim, 1)
        for B in batches:
            qs = torch.randn((8, B, dim), device="cuda")
            out = eng.search(qs[0], 10, metric)
            i = [0]

            def step():
                eng.search(qs[i[0] % 8], 10, metric, out=out)
                i[0] += 1
            ms = min(timed(step, 200 if B <= 64 else 5) for _ in range(3))   # best of 3 runs of 200 searches
            eng.set_kernel_timing(True)
            kms, cands = [], []
            for _ in range(5):
                step()
                st = eng.stats()
                kms.append(st["last_kernel_ms"])
                cands.append(st["last_candidates"] / B)
            eng.set_kernel_timing(False)
            kernel_ms = sum(kms) / len(kms)
            shadow = n * ((dim + 7) // 8 * 8) * 2 + 8 * n
            r = {"rows": n, "dim": dim, "metric": metric, "batch": B, "step_ms": ms, "qps": B / ms * 1e3,
                 "kernel": st["last_kernel"], "filter_ms": kernel_ms, "around_filter_us": (ms - kernel_ms) * 1e3,
                 "filter_gbs": shadow / kernel_ms / 1e6, "filter_frac_hbm": shadow / kernel_ms / 1e6 / PEAK,
                 "useful_tflops": 2.0 * n * dim * B / kernel_ms / 1e9, "candidates_per_query": sum(cands) / len(cands)}
            res.append(r)
            print(json.dumps(r), flush=True)
        eng.close()
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    (ROOT / "gpurun_out" / "filter_probe.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
