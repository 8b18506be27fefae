"""Measure every BASELINE.json config on ONE GPU (device-resident queries, CUDA events) and write
gpurun_out/config_sweep.json.  C4 is its per-GPU shard (12.5M of the 100M rows)."""
import json, os, sys, time, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import numpy as np, torch
import wdbx_b200

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0

def fill(eng, n, dim, seed):
    eng.reserve(0, n)
    g = torch.Generator(device="cuda").manual_seed(seed)
    done = 0
    while done < n:
        m = min(1 << 20, n - done); eng.append(0, torch.randn((m, dim), generator=g, device="cuda")); done += m

def timed(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

res = []
BF16_PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("bf16_tflops", 1661.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 1590.0

def run(name, n, dim, dtype, metric, k, B, iters, gemm=16, shadow_mb=None):
    os.environ["WDBX_B200_GEMM_MIN_BATCH"] = str(gemm)
    if shadow_mb is None: os.environ.pop("WDBX_B200_SHADOW_MIN_MB", None)
    else: os.environ["WDBX_B200_SHADOW_MIN_MB"] = str(shadow_mb)
    eng = wdbx_b200.Engine(0, dim, dtype, 1)
    fill(eng, n, dim, 1)
    qs = torch.randn((8, B, dim), device="cuda")
    out = eng.search(qs[0], k, metric)
    i = [0]
    def step():
        eng.search(qs[i[0] % 8], k, metric, out=out); i[0] += 1
    ms = timed(step, iters)
    eng.set_kernel_timing(True)
    kms = []
    for _ in range(3):
        step(); st = eng.stats(); kms.append(st["last_kernel_ms"])
    eng.set_kernel_timing(False)
    kid, kernel_ms = st["last_kernel"], sum(kms) / len(kms)
    eb = 2 if dtype == "bf16" else 4
    stored = n * dim * eb + (4 * n if metric == "cosine" else 0)
    r = {"config": name, "rows": n, "dim": dim, "dtype": dtype, "metric": metric, "k": k, "batch": B,
         "ms_per_call": ms, "qps": B / ms * 1e3, "dominant_kernel_ms": kernel_ms}
    if kid == 1:
        passes = (B + 7) // 8
        r.update(kernel="K1 scan_topk (stored rows)", kernel_hbm_gbs=stored * passes / kernel_ms / 1e6,
                 frac_of_measured_hbm_peak=stored * passes / kernel_ms / 1e6 / PEAK)
    else:
        if kid == 3:   # small-batch kernel over the 1-byte shadow (+ 1/|x|, scale, |r| per row)
            shadow = n * ((dim + 15) // 16 * 16) + 12 * n
            kname = "K2b gemm_filter_small (kind::i8 tcgen05 over the 1-byte shadow, in-kernel exact re-score, one launch)"
        else:
            shadow = n * ((dim + 7) // 8 * 8) * 2 + 4 * n
            kname = "K2b gemm_filter (bf16 tcgen05 over the 2-byte shadow) + exact fp32 refine"
        fl = 2.0 * n * dim * B
        r.update(kernel=kname,
                 kernel_hbm_gbs=shadow / kernel_ms / 1e6, frac_of_measured_hbm_peak=shadow / kernel_ms / 1e6 / PEAK,
                 useful_tflops=fl / ms / 1e9, frac_of_measured_bf16_peak=fl / ms / 1e9 / BF16_PEAK,
                 fp32_scan_equivalent_gbs=stored / ms / 1e6)
    print(json.dumps(r), flush=True)
    res.append(r)
    eng.close()

# C1 through the public API (host lists in, tuples out)
tmp = tempfile.mkdtemp()
db = wdbx_b200.WDBX(vector_dimension=384, num_shards=2, data_dir=tmp, log_level="WARNING")
X = np.random.default_rng(0).standard_normal((10000, 384)).astype(np.float32)
db.vector_store.batch_store({f"d{i}": X[i] for i in range(10000)})
ql = [np.random.default_rng(i).standard_normal(384).astype(np.float32).tolist() for i in range(32)]
for q in ql[:5]: db.vector_search(q, limit=5)
t0 = time.perf_counter()
for j in range(200): db.vector_search(ql[j % 32], limit=5)
us = (time.perf_counter() - t0) / 200 * 1e6
r = {"config": "C1 quick-start 10k x 384, 2 shards, k=5, public API end to end", "us_per_query": us, "qps": 1e6 / us,
     "engine_device_ms": db.get_stats()["gpu"]["engine"]["last_search_ms"]}
print(json.dumps(r), flush=True); res.append(r)
db.close()

run("C2 1M x 384 fp32 cosine B=1", 1_000_000, 384, "fp32", "cosine", 10, 1, 50)
run("C2 1M x 384 fp32 cosine B=1, filter path forced", 1_000_000, 384, "fp32", "cosine", 10, 1, 50, shadow_mb=0)
run("C3 10M x 768 fp32 cosine B=1", 10_000_000, 768, "fp32", "cosine", 10, 1, 20)
run("C3 10M x 768 fp32 cosine B=1, K1 scan of the stored rows forced", 10_000_000, 768, "fp32", "cosine", 10, 1, 20, shadow_mb=-1)
run("C3 10M x 768 fp32 cosine B=8", 10_000_000, 768, "fp32", "cosine", 10, 8, 10)
run("C3 10M x 768 fp32 cosine B=8, K1 (8 queries per pass) forced", 10_000_000, 768, "fp32", "cosine", 10, 8, 10, shadow_mb=-1)
run("C3 10M x 768 fp32 cosine B=64", 10_000_000, 768, "fp32", "cosine", 10, 64, 5)
run("C3 10M x 768 fp32 cosine B=256", 10_000_000, 768, "fp32", "cosine", 10, 256, 5)
run("C3 10M x 768 fp32 cosine B=1024", 10_000_000, 768, "fp32", "cosine", 10, 1024, 3)
run("C4 shard 12.5M x 384 bf16 ip k=100 B=1", 12_500_000, 384, "bf16", "ip", 100, 1, 20)
run("C5 5M x 1536 fp32 l2 B=4096", 5_000_000, 1536, "fp32", "l2", 10, 4096, 2)
run("C5 5M x 1536 fp32 l2 B=1", 5_000_000, 1536, "fp32", "l2", 10, 1, 20)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "config_sweep.json").write_text(json.dumps(res, indent=1))
