#!/usr/bin/env python
"""bench.py -- exact top-10 queries/sec of the WDBX vector_search hot path on B200.

Workload (BASELINE.json metric / configs[2], "C3"): 10M x 768 fp32 cosine, top-10, query batch 1,
rows striped over the N GPUs of one box (strong scaling: the total matrix is fixed).
A "step" is one query over the whole matrix: every rank filters its rows with the bf16 tensor-core
filter over the 2-byte shadow of its fp32 rows (K2b: half the HBM bytes of a scan of the stored rows,
rigorous error bound), re-scores the few hundred surviving rows from the fp32 rows with the streaming
kernel's exact arithmetic (refine: results bit-identical to the fp32 scan K1), and the ranks exchange
their packed top-10 keys over NVLink peer memory and merge them on the device.

One JSON line on rank 0 (see the keys in main()).  `value` is measured with queries resident in
HBM, `e2e` through the public host API (VectorStore.search: host query in, (id, score, metadata)
tuples out, H2D + D2H inside the timed region).  `--impl reference` times the reference's CPU
path (oracle/flat_ip.c restatement of FAISS-Flat, all host threads) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))

N_ROWS = 10_000_000
DIM = 768
K = 10
METRIC = "cosine"
CHUNK = 1_000_000          # generation chunk (global rows); a multiple of every supported N
N_QUERIES = 64             # distinct queries cycled through the steps
CPU_SAMPLE_ROWS = 1_000_000
METRIC_NAME = "exact top-10 queries/sec, 10Mx768 fp32 cosine, batch 1"
WORKLOAD = "C3: 10M x 768 fp32 cosine, top-10, query batch 1, rows striped over N GPUs"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic_from_profile(kernel_id: int):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    name = "gemm_filter_b1_ncu_summary.json" if kernel_id == 2 else "scan_topk_c3_ncu_summary.json"
    try:
        return float(json.loads((ROOT / "profiles" / name).read_text()).get("dram_bytes_per_launch")), f"profiles/{name}"
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.active = False
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self._nv = None

    def run(self):
        if not self.ok:
            return
        nv = self._nv
        while not self._stop_evt.is_set():
            if self.active:
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_scan_qps(steps: int, warmup: int, rows: int = CPU_SAMPLE_ROWS, X=None, Q=None):
    """QPS of the reference's CPU path (C restatement of FAISS-Flat: normalised rows . normalised
    query, heap top-k, all OpenMP threads) on `rows` rows, and the thread count used."""
    import ctypes as C

    import numpy as np

    import __graft_entry__ as ge

    lib = C.CDLL(str(ge.build_oracle()))
    lib.oracle_flat_search.restype = C.c_int
    lib.oracle_flat_search.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_normalize_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    lib.oracle_num_threads.restype = C.c_int
    cores = len(os.sched_getaffinity(0))
    if X is None:
        rng = np.random.default_rng(1234)
        X = rng.standard_normal((rows, DIM), dtype=np.float32)
    if Q is None:
        Q = np.random.default_rng(4321).standard_normal((N_QUERIES, DIM), dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    lib.oracle_normalize_rows(X.ctypes.data, X.shape[0], DIM)       # FaissIndex.add normalises at ingest
    out_r = np.empty(K, np.int64)
    out_s = np.empty(K, np.float32)
    results = []

    def one(i):
        q = Q[i % Q.shape[0]]
        nrm = np.linalg.norm(q)
        qn = (q / nrm if nrm > 0 else q).astype(np.float32)        # FaissIndex.search normalises the query
        n = lib.oracle_flat_search(X.ctypes.data, X.shape[0], DIM, qn.ctypes.data, 0, K, None, out_r.ctypes.data,
                                   out_s.ctypes.data, cores)   # explicit: torchrun exports OMP_NUM_THREADS=1
        assert n == min(K, X.shape[0])
        return out_r.copy(), out_s.copy()

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(steps):
        results.append(one(i))
    dt = time.perf_counter() - t0
    return steps / dt, cores, dt, results


def run_reference(args):
    """--impl reference: the reference's CPU search path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 40)
    warmup = min(args.warmup, 3)
    qps_sample, cores, dt, _ = _cpu_scan_qps(steps, warmup)
    scale = CPU_SAMPLE_ROWS / N_ROWS
    value = qps_sample * scale
    sample = (f"{CPU_SAMPLE_ROWS} of {N_ROWS} rows x {DIM} (same distribution), {steps} single-query steps, "
              f"QPS scaled by {scale:g} (the scan is O(N*D))")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": N_ROWS, "dim": DIM, "k": K, "metric": METRIC, "batch": 1},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    import wdbx_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)")
    if rank == 0:
        ge.build_cuda()
    torch.cuda.set_device(local)
    ctx = wdbx_b200.DistContext.from_env(local)
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local)
    n_rows, steps, warmup = args.rows, args.steps, args.warmup

    import tempfile

    tmp = tempfile.mkdtemp(prefix="wdbx_b200_bench_")
    store = wdbx_b200.VectorStore(DIM, tmp, num_shards=1, config=wdbx_b200.WDBXConfig(
        {"GPU_METRIC": METRIC, "GPU_STRICT": True}), dist=ctx)
    store.engine.reserve(0, store.shard_map.local_count(n_rows, rank))

    # ---- synthetic data: iid N(0,1) fp32, generated on device chunk by chunk from per-chunk seeds so
    # the global matrix is identical for every N; rank r keeps global rows r, r+N, ...
    cpu_prefix = []
    for c in range((n_rows + CHUNK - 1) // CHUNK):
        m = min(CHUNK, n_rows - c * CHUNK)
        g = torch.Generator(device=dev).manual_seed(1234 + 1000 * 3 + c)
        x = torch.randn((m, DIM), generator=g, device=dev, dtype=torch.float32)
        if rank == 0 and c * CHUNK < CPU_SAMPLE_ROWS and not args.no_cpu_baseline:
            take = min(m, CPU_SAMPLE_ROWS - c * CHUNK)
            cpu_prefix.append(x[:take].cpu().numpy())
        mine = x[rank::world].contiguous() if world > 1 else x
        store.bulk_load({"local": mine, "total": m}, id_prefix=f"c{c}_")
        del x, mine
    gq = torch.Generator(device=dev).manual_seed(4321 + 3)
    Qd = torch.randn((N_QUERIES, DIM), generator=gq, device=dev, dtype=torch.float32)
    Qh = Qd.cpu().numpy()
    Qlists = [Qh[i].tolist() for i in range(N_QUERIES)]
    torch.cuda.synchronize()
    local_rows = store.engine.stats()["rows_total"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = store.engine.stats()["kernel_launches"]

    # ---- value: device-resident queries, K steps, CUDA events on the launching (current) stream
    qs = [Qd[i:i + 1] for i in range(N_QUERIES)]
    for i in range(max(warmup, 3)):
        store.search_device(qs[i % N_QUERIES], K)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    l_before = store.engine.stats()["kernel_launches"]
    e0.record()
    for i in range(steps):
        out = store.search_device(qs[i % N_QUERIES], K)
    e1.record()
    barrier()
    sampler.active = False
    gpu_launches = store.engine.stats()["kernel_launches"] - l_before
    ms_total = reduce_max(e0.elapsed_time(e1))
    ms_per_step = ms_total / steps
    qps = 1e3 / ms_per_step

    # ---- roofline of the dominant kernel: the engine brackets ITS launches with CUDA events on the search
    # stream (wdbx_b200_set_kernel_timing); which kernel that is (K2b filter over the bf16 shadow, or the K1
    # scan of the stored rows) is reported by the engine, not assumed here
    store.engine.set_kernel_timing(True)
    kout = store.engine.search(qs[0], K, METRIC)
    for i in range(3):
        store.engine.search(qs[i], K, METRIC, out=kout)
    n_k = max(10, min(steps, 100))
    k_ms, kernel_id = [], 0
    for i in range(n_k):
        store.engine.search(qs[i % N_QUERIES], K, METRIC, out=kout)
        st = store.engine.stats()          # synchronises on the kernel's end event
        k_ms.append(st["last_kernel_ms"])
        kernel_id = st["last_kernel"]
    store.engine.set_kernel_timing(False)
    kernel_ms = reduce_max(sum(k_ms) / len(k_ms))
    ld16 = (DIM + 7) // 8 * 8
    if kernel_id == 2:
        kernel_name = "gemm_filter_small_kernel (K2b bf16 tcgen05 filter, B<=16 variant, over the 2-byte shadow rows)"
        algo_bytes = local_rows * ld16 * 2 + local_rows * 4     # 2-byte shadow rows + 1/|x| per row
        algo_note = "rows x dim x 2 (bf16 shadow) + rows x 4 (1/|x|): the bytes THIS kernel must read"
    else:
        kernel_name = "scan_topk_kernel (K1)"
        algo_bytes = local_rows * DIM * 4 + local_rows * 4      # rows + 1/|x| per row
        algo_note = "rows x dim x 4 + rows x 4 (SURVEY.md 8d)"
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    k1_bytes = local_rows * DIM * 4 + local_rows * 4
    peak, peak_src = _peaks()
    traffic, traffic_src = _traffic_from_profile(kernel_id)   # ncu --set full capture on the 10M-row matrix (N=1)
    if traffic is not None:
        traffic_src += " (dram read+write per launch, 10M rows on one GPU)"
        if local_rows != N_ROWS:
            traffic = traffic * local_rows / N_ROWS
            traffic_src += f", scaled by rows_per_gpu/{N_ROWS}"

    # ---- e2e: public host API (list of floats in, tuples out), H2D + D2H + id mapping inside
    n_e2e = max(10, min(steps, 100))
    for i in range(3):
        store.search(Qlists[i], limit=K)
    barrier()
    step_ms = []
    t0 = time.perf_counter()
    for i in range(n_e2e):
        t1 = time.perf_counter()
        res = store.search(Qlists[i % N_QUERIES], limit=K)
        step_ms.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_qps = n_e2e / e2e_s
    assert len(res) == K
    # where the end-to-end time goes (informational): C-ABI host call alone vs the Python facade
    e2e_detail = {"device_ms_last_call": store.engine.stats().get("last_search_ms"),
                  "step_ms_median": statistics.median(step_ms), "step_ms_max": max(step_ms),
                  "step_ms_first5": [round(x, 3) for x in step_ms[:5]]}
    if world == 1:
        t0 = time.perf_counter()
        for i in range(20):
            store.engine.search_host(Qh[i % N_QUERIES], K, metric=METRIC)
        e2e_detail["c_abi_search_host_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        e2e_detail["device_ms_last_call"] = store.engine.stats().get("last_search_ms")
    sampler.stop()

    # ---- parity gate on the timed data: exact fp64 re-score of the returned rows' neighbourhood
    parity = None
    if rank == 0 or world > 1:
        parity = _parity_check(store, Qd, out_check_queries=2, world=world, rank=rank, dev=dev, n_rows=n_rows)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        Xs = np.concatenate(cpu_prefix) if cpu_prefix else None
        sample_rows = Xs.shape[0]
        qps_s, cores, dt, cres = _cpu_scan_qps(12, 2, rows=sample_rows, X=Xs, Q=Qh)
        scale = sample_rows / n_rows
        cpu = {"value": qps_s * scale, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_rows} of {n_rows} rows of the same matrix, 12 single-query steps "
                         f"({dt:.1f} s), QPS scaled by {scale:g}"}

    if rank == 0:
        line = {
            "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if n_rows == N_ROWS else f"REDUCED {n_rows} x {DIM} (not the named config)",
                       "rows": n_rows, "dim": DIM, "k": K, "metric": METRIC, "batch": 1,
                       "rows_per_gpu": local_rows, "parallelism": f"row-striped x{world}",
                       "l2_policy": "inputs larger than L2 (>=3.8 GB per GPU streamed per step vs 126 MB L2); "
                                    "64 distinct queries cycled"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo_bytes,
                         "algorithmic_bytes_definition": algo_note, "peak_source": peak_src,
                         # the same launch expressed in SURVEY.md 8d's K1 bytes (fp32 rows): what a scan of the
                         # stored rows would have had to stream in this time
                         "fp32_scan_equivalent_gbs": k1_bytes / (kernel_ms * 1e-3) / 1e9,
                         "note": ("achieved/frac use the bytes this kernel must read; with SURVEY.md 8d's K1 bytes "
                                  "(fp32 rows, which this path no longer streams) the same launch would read as "
                                  f"frac {k1_bytes / (kernel_ms * 1e-3) / 1e9 / peak:.2f}") if kernel_id == 2 else None},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": DIM * 4,
                    "d2h_bytes_per_step": K * 20 + 4, "api": "VectorStore.search (host list in, tuples out)",
                    "steps": n_e2e, "ms_per_step": e2e_s / n_e2e * 1e3, "detail": e2e_detail},
            "gpu_launches": int(gpu_launches),
            "clocks": sampler.summary(),
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    store.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _parity_check(store, Qd, out_check_queries, world, rank, dev, n_rows):
    """Checker only (not the product path): torch fp32 chunked scores on this rank's rows -> global
    top-k via all-gather on the host, compared with the engine's ids."""
    import torch
    import torch.distributed as dist

    ok = True
    worst = 0.0
    for b in range(out_check_queries):
        q = Qd[b:b + 1]
        got = store.search_device(q, K)
        got_g = got["gids"][0].cpu()
        got_s = got["scores"][0].cpu()
        # device rows are not exposed as tensors; re-generate this rank's stripe chunk by chunk
        best_s, best_g = [], []
        for c in range((n_rows + CHUNK - 1) // CHUNK):
            m = min(CHUNK, n_rows - c * CHUNK)
            g = torch.Generator(device=dev).manual_seed(1234 + 1000 * 3 + c)
            x = torch.randn((m, DIM), generator=g, device=dev, dtype=torch.float32)
            xs = x[rank::world] if world > 1 else x
            s = (xs.double() @ q[0].double()) / (xs.double().norm(dim=1) * q[0].double().norm())
            v, i = torch.topk(s, min(K, s.numel()))
            best_s.append(v)
            best_g.append(c * CHUNK + rank + i * world)
            del x, xs, s
        v = torch.cat(best_s)
        gsel = torch.cat(best_g)
        if world > 1:
            vs = [torch.empty_like(v) for _ in range(world)]
            gs = [torch.empty_like(gsel) for _ in range(world)]
            dist.all_gather(vs, v)
            dist.all_gather(gs, gsel)
            v, gsel = torch.cat(vs), torch.cat(gs)
        top = torch.topk(v, K)
        want_g = gsel[top.indices].cpu()
        want_s = top.values.cpu()
        ok = ok and bool((want_g == got_g).all())
        worst = max(worst, float((got_s.double() - want_s).abs().max()))
    return {"checked_queries": out_check_queries, "ids_match_fp64_checker": ok, "max_abs_score_err": worst}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="debug only: a reduced matrix is flagged in config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
