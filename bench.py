#!/usr/bin/env python
"""bench.py -- exact top-10 queries/sec of the WDBX vector_search hot path on B200.

Workload (BASELINE.json metric / configs[2], "C3"): 10M x 768 fp32 cosine, top-10, query batch 1,
rows striped over the N GPUs of one box (strong scaling: the total matrix is fixed).
A "step" is one query over the whole matrix: every rank filters its rows with the tensor-core filter
over the 1-byte (int8) shadow of its fp32 rows (K2b small-batch kernel: a quarter of the HBM bytes of a scan of
the stored rows, rigorous data-derived error bound), re-scores the few hundred surviving rows from the fp32
rows with the streaming kernel's exact arithmetic inside the same launch (results bit-identical to the fp32
scan K1), and the ranks exchange their packed top-10 keys over NVLink peer memory and merge them on the device.

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
CPU_SAMPLE_ROWS = 1_000_000    # rows of the cpu_baseline leg inside the GPU arm (bounded: ~1 s of CPU work)
PARITY_QUERIES = 16
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
    names = (["r02_gemm_filter_small_i8_c3_ncu_summary.json"] if kernel_id == 3
             else ["r02_gemm_filter_small_c3_ncu_summary.json", "gemm_filter_b1_ncu_summary.json"] if kernel_id == 2
             else ["scan_topk_c3_ncu_summary.json"])
    for name in names:
        try:
            return float(json.loads((ROOT / "profiles" / name).read_text()).get("dram_bytes_per_launch")), f"profiles/{name}"
        except Exception:
            continue
    return None, None


# ------------------------------------------------------------------------------------------ synthetic data
def _chunk_seed(c: int) -> int:
    return 1234 + 1000 * 3 + c


def _gen_chunk(dev, c: int, m: int):
    """Chunk c (rows c*CHUNK ...) of the benchmark matrix: iid N(0,1) fp32 from a per-chunk seed, generated on the
    GPU -- BOTH arms call this, so the reference arm scans exactly the matrix the GPU arm holds."""
    import torch

    g = torch.Generator(device=dev).manual_seed(_chunk_seed(c))
    return torch.randn((m, DIM), generator=g, device=dev, dtype=torch.float32)


def _gen_queries(dev):
    import torch

    gq = torch.Generator(device=dev).manual_seed(4321 + 3)
    return torch.randn((N_QUERIES, DIM), generator=gq, device=dev, dtype=torch.float32)


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


def _host_ram_bytes() -> int:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    try:
        return os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        return 0


def run_reference(args):
    """--impl reference: the reference's CPU search path (oracle/flat_ip.c, all host threads) on the box's host
    cores, over the SAME seeded matrix and queries as the GPU arm, exactly --steps / --warmup steps.  The full
    10M x 768 matrix (30.7 GB) when the host has >= 48 GB available, else a prefix sample with QPS scaled."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    steps, warmup = args.steps, args.warmup
    n_rows = args.rows
    full = _host_ram_bytes() >= 48 * (1 << 30) and not args.cpu_sample
    rows = n_rows if full else min(n_rows, CPU_SAMPLE_ROWS)
    same_data = False
    X = np.empty((rows, DIM), dtype=np.float32)
    try:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device for the shared generator")
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        for c in range((rows + CHUNK - 1) // CHUNK):
            m = min(CHUNK, n_rows - c * CHUNK)          # generate the chunk exactly as the GPU arm does ...
            take = min(m, rows - c * CHUNK)             # ... and keep the part inside the sample
            X[c * CHUNK: c * CHUNK + take] = _gen_chunk(dev, c, m)[:take].cpu().numpy()
        Q = _gen_queries(dev).cpu().numpy()
        same_data = True
        del dev
        torch.cuda.empty_cache()
    except Exception:
        rng = np.random.default_rng(1234)
        for c in range((rows + CHUNK - 1) // CHUNK):
            take = min(CHUNK, rows - c * CHUNK)
            X[c * CHUNK: c * CHUNK + take] = rng.standard_normal((take, DIM), dtype=np.float32)
        Q = np.random.default_rng(4321).standard_normal((N_QUERIES, DIM), dtype=np.float32)
    qps_sample, cores, dt, _ = _cpu_scan_qps(steps, warmup, rows=rows, X=X, Q=Q)
    scale = rows / n_rows
    value = qps_sample * scale
    if full:
        sample = f"the full {n_rows} x {DIM} matrix, {steps} single-query steps ({dt:.1f} s)"
    else:
        sample = (f"first {rows} of {n_rows} rows x {DIM}, {steps} single-query steps ({dt:.1f} s), QPS scaled by "
                  f"{scale:g} (the scan is O(N*D)); host RAM available {_host_ram_bytes() / 2**30:.0f} GiB < 48 GiB"
                  if not args.cpu_sample else f"first {rows} of {n_rows} rows (--cpu-sample), QPS scaled by {scale:g}")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic" + (" (the GPU arm's seeded matrix and queries)" if same_data else " (numpy generator: no CUDA device)"),
        "config": {"workload": WORKLOAD if n_rows == N_ROWS else f"REDUCED {n_rows} x {DIM} (not the named config)",
                   "rows": n_rows, "dim": DIM, "k": K, "metric": METRIC, "batch": 1,
                   "rows_per_gpu": n_rows // max(args.gpus, 1), "parallelism": f"row-striped x{args.gpus}",
                   "l2_policy": "host arm: the matrix (30.7 GB) is far larger than any CPU cache; 64 distinct queries cycled"},
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
        x = _gen_chunk(dev, c, m)
        if rank == 0 and c * CHUNK < CPU_SAMPLE_ROWS and not args.no_cpu_baseline:
            take = min(m, CPU_SAMPLE_ROWS - c * CHUNK)
            cpu_prefix.append(x[:take].cpu().numpy())
        mine = x[rank::world].contiguous() if world > 1 else x
        store.bulk_load({"local": mine, "total": m}, id_prefix=f"c{c}_")
        del x, mine
    Qd = _gen_queries(dev)
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

    # ---- value: device-resident queries, K steps, CUDA events on the launching (current) stream.  All queries are
    # resident before the timed region, which is exactly the contract of the engine's `overlap` option: the next
    # search streams its first tiles while the previous one finishes its tail (results stay in launch order)
    store.engine.set_option("overlap", 1)
    qs = [Qd[i:i + 1] for i in range(N_QUERIES)]
    warmup = max(warmup, 3)
    for i in range(warmup):
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
    ld8 = (DIM + 15) // 16 * 16
    if kernel_id == 3:
        kernel_name = ("gemm_filter_small_kernel<int8> (K2b, B<=16: kind::i8 tcgen05 filter over the 1-byte shadow rows + in-kernel "
                       "exact re-score, last-CTA merge and cross-GPU exchange -- the whole search is this one launch)")
        algo_bytes = local_rows * ld8 + local_rows * 12         # 1-byte shadow rows + 1/|x|, row scale, residual norm
        algo_note = "rows x dim x 1 (int8 shadow) + rows x 12 (1/|x|, scale, |r|): the bytes THIS kernel must read"
    elif kernel_id == 2:
        kernel_name = ("gemm_filter_small_kernel (K2b, B<=16: bf16 tcgen05 filter over the 2-byte shadow rows + in-kernel exact "
                       "re-score, last-CTA merge and cross-GPU exchange -- the whole search is this one launch)")
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
        traffic_src += " (ncu --set full capture, not this run: dram read+write per launch, 10M rows on one GPU)"
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

    # ---- parity gate on the timed data: fp64 checker over the whole matrix, PARITY_QUERIES queries, through every
    # route the store can take at this N (device exchange / NCCL all-gather + merge kernel)
    checker = _Fp64Checker(Qd, world, rank, dev, n_rows)
    routes = {"default": lambda q: store.search_device(q, K)}
    if world > 1:
        routes["nccl_allgather_merge"] = lambda q: store.engine.merge(
            store.dist.all_gather_keys(store.engine.search(q, K, METRIC)["keys"]))
    parity = {"checked_queries": PARITY_QUERIES}
    for name, fn in routes.items():
        parity[name] = checker.check(fn, range(PARITY_QUERIES))
    parity["ids_match"] = all(v["ids_match_fp64_checker"] for v in parity.values() if isinstance(v, dict))

    # ---- extra (not the headline): the other regimes on the same resident matrix, each with its own parity flag
    extra = None if args.no_extra else _extra_regimes(store, Qd, qs, checker, local_rows, n_rows, world, dev, barrier,
                                                      reduce_max, peak)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        Xs = np.concatenate(cpu_prefix) if cpu_prefix else None
        sample_rows = Xs.shape[0]
        qps_s, cores, dt, cres = _cpu_scan_qps(40, 3, rows=sample_rows, X=Xs, Q=Qh)
        scale = sample_rows / n_rows
        cpu = {"value": qps_s * scale, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_rows} of {n_rows} rows of the same matrix, 40 single-query steps "
                         f"({dt:.1f} s), QPS scaled by {scale:g}"}

    if rank == 0:
        line = {
            "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if n_rows == N_ROWS else f"REDUCED {n_rows} x {DIM} (not the named config)",
                       "rows": n_rows, "dim": DIM, "k": K, "metric": METRIC, "batch": 1,
                       "rows_per_gpu": local_rows, "parallelism": f"row-striped x{world}",
                       "overlap": "value: consecutive device-resident searches overlap on the device (engine option); "
                                  "e2e: one search at a time",
                       "l2_policy": "inputs larger than L2 (>=3.8 GB per GPU streamed per step vs 126 MB L2); "
                                    "64 distinct queries cycled"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # the same launch in SURVEY.md 8d's bytes (fp32 rows + 1/|x|): > 1 means "faster than any scan of
                         # the stored fp32 rows could be" -- the gain is algorithmic (2-byte shadow), not bandwidth
                         "frac_8d_bytes": k1_bytes / (kernel_ms * 1e-3) / 1e9 / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo_bytes,
                         "algorithmic_bytes_definition": algo_note, "peak_source": peak_src,
                         # the same launch expressed in SURVEY.md 8d's K1 bytes (fp32 rows): what a scan of the
                         # stored rows would have had to stream in this time
                         "fp32_scan_equivalent_gbs": k1_bytes / (kernel_ms * 1e-3) / 1e9,
                         "step_vs_kernel": ("kernel_ms is one launch timed ALONE (events around it serialise the stream); in the "
                                            "timed region consecutive searches overlap on the device -- search n+1 streams its "
                                            "first tiles while search n re-scores / merges / exchanges -- so ms_per_step can be "
                                            "smaller than kernel_ms"),
                         "note": ("achieved/frac use the bytes this kernel must read; with SURVEY.md 8d's K1 bytes "
                                  "(fp32 rows, which this path no longer streams) the same launch would read as "
                                  f"frac {k1_bytes / (kernel_ms * 1e-3) / 1e9 / peak:.2f}") if kernel_id >= 2 else None},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": DIM * 4,
                    "d2h_bytes_per_step": K * 20 + 4, "api": "VectorStore.search (host list in, tuples out)",
                    "steps": n_e2e, "ms_per_step": e2e_s / n_e2e * 1e3, "detail": e2e_detail},
            "gpu_launches": int(gpu_launches),
            "clocks": sampler.summary(),
            "parity": parity,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    store.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _Fp64Checker:
    """Checker only (never the product path): exact fp64 cosine of EVERY row against a set of queries (chunk by chunk,
    the rows are re-generated from their seeds), global top-k via all-gather; cached per query index."""

    def __init__(self, Qd, world, rank, dev, n_rows):
        self.Qd, self.world, self.rank, self.dev, self.n_rows = Qd, world, rank, dev, n_rows
        self.want = {}

    def _compute(self, idx):
        import torch
        import torch.distributed as dist

        idx = [i for i in idx if i not in self.want]
        if not idx:
            return
        q = self.Qd[idx].double()
        qn = q / q.norm(dim=1, keepdim=True)
        best_s, best_g = [], []
        for c in range((self.n_rows + CHUNK - 1) // CHUNK):
            m = min(CHUNK, self.n_rows - c * CHUNK)
            x = _gen_chunk(self.dev, c, m)
            xs = (x[self.rank::self.world] if self.world > 1 else x).double()
            s = (xs @ qn.T) / xs.norm(dim=1, keepdim=True)          # [rows, nq]
            v, i = torch.topk(s, min(K, s.shape[0]), dim=0)
            best_s.append(v)
            best_g.append(c * CHUNK + self.rank + i * self.world)
            del x, xs, s
        v, g = torch.cat(best_s), torch.cat(best_g)
        if self.world > 1:
            vs = [torch.empty_like(v) for _ in range(self.world)]
            gs = [torch.empty_like(g) for _ in range(self.world)]
            dist.all_gather(vs, v)
            dist.all_gather(gs, g)
            v, g = torch.cat(vs), torch.cat(gs)
        top = torch.topk(v, K, dim=0)
        for j, i in enumerate(idx):
            self.want[i] = (g[:, j][top.indices[:, j]].cpu(), top.values[:, j].cpu())

    def check(self, search_fn, idx, batch=None):
        """search_fn(q [B, dim]) -> result dict; idx: query indices (rows of Qd) or, with `batch`, rows of `batch`
        whose first len(idx) rows are Qd[idx]."""
        import torch

        idx = list(idx)
        self._compute(idx)
        ok, worst = True, 0.0
        if batch is not None:
            out = search_fn(batch)
            got = [(out["gids"][j].cpu(), out["scores"][j].cpu()) for j in range(len(idx))]
        else:
            got = []
            for i in idx:
                out = search_fn(self.Qd[i:i + 1])
                got.append((out["gids"][0].cpu(), out["scores"][0].cpu()))
        torch.cuda.synchronize()
        for i, (gg, gs) in zip(idx, got):
            wg, ws = self.want[i]
            ok = ok and bool((wg == gg).all())
            worst = max(worst, float((gs.double() - ws).abs().max()))
        return {"ids_match_fp64_checker": ok, "max_abs_score_err": worst, "queries": len(idx)}


def _tf32_peak_tflops(dev):
    """Measured dense TF32 throughput (torch.matmul, 8192^3, best of 5): the denominator for the 3xTF32 kernel K2."""
    import torch

    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((8192, 8192), device=dev)
        b = torch.randn((8192, 8192), device=dev)
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def _extra_regimes(store, Qd, qs, checker, local_rows, n_rows, world, dev, barrier, reduce_max, peak):
    """(i) the fp32 streaming scan K1 forced on the same matrix -- SURVEY.md 8d's bytes, the north star's ">= 80 % of
    the HBM roofline" claim; (ii) C3 with a batch of 1024 queries (tensor-core regime) as useful TFLOP/s against the
    measured bf16 peaks.  Device-timed like `value`; every number carries its own parity flag."""
    import torch

    extra = {}
    eng = store.engine
    # ---- (o) the headline loop again with consecutive searches fully ORDERED (engine option overlap = 0, the library
    # default): what one stream of dependent searches sees
    try:
        eng.set_option("overlap", 0)
        for i in range(3):
            store.search_device(qs[i], K)
        barrier()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            store.search_device(qs[i % N_QUERIES], K)
        e1.record()
        barrier()
        ms = reduce_max(e0.elapsed_time(e1)) / n
        extra["no_overlap"] = {"ms_per_step": ms, "queries_per_s": 1e3 / ms, "steps": n,
                               "note": "same device-resident loop as `value`, consecutive searches ordered on the stream"}
    except Exception as e:
        extra["no_overlap"] = {"error": repr(e)}
    finally:
        eng.set_option("overlap", 1)
    # ---- (i) K1 forced
    try:
        eng.set_option("shadow_min_mb", -1)
        for i in range(3):
            store.search_device(qs[i], K)
        barrier()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            store.search_device(qs[i % N_QUERIES], K)
        e1.record()
        barrier()
        step_ms = reduce_max(e0.elapsed_time(e1)) / n
        eng.set_kernel_timing(True)
        kms = []
        kout = eng.search(qs[0], K, METRIC)
        for i in range(10):
            eng.search(qs[i], K, METRIC, out=kout)
            st = eng.stats()
            kms.append(st["last_kernel_ms"])
            kid = st["last_kernel"]
        eng.set_kernel_timing(False)
        kernel_ms = reduce_max(sum(kms) / len(kms))
        bytes_8d = local_rows * DIM * 4 + local_rows * 4
        par = checker.check(lambda q: store.search_device(q, K), range(PARITY_QUERIES))
        extra["k1_fp32_scan_forced"] = {
            "kernel": "scan_topk_kernel (K1)" if kid == 1 else f"unexpected kernel id {kid}", "ms_per_step": step_ms,
            "queries_per_s": 1e3 / step_ms, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_8d,
            "achieved_gbs": bytes_8d / (kernel_ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_8d / (kernel_ms * 1e-3) / 1e9 / peak,
            "bytes_definition": "rows x dim x 4 + rows x 4 (SURVEY.md 8d)", "parity": par}
    except Exception as e:  # never lose the headline line to an extra
        extra["k1_fp32_scan_forced"] = {"error": repr(e)}
    finally:
        eng.set_option("shadow_min_mb", 1024)
    # ---- (ii) batch of 1024 queries
    try:
        B = 1024
        g = torch.Generator(device=dev).manual_seed(99)
        Qb = torch.randn((B, DIM), generator=g, device=dev, dtype=torch.float32)
        Qb[:PARITY_QUERIES] = Qd[:PARITY_QUERIES]
        for _ in range(2):
            store.search_device(Qb, K)
        barrier()
        n = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            store.search_device(Qb, K)
        e1.record()
        barrier()
        ms = reduce_max(e0.elapsed_time(e1)) / n
        tflops = 2.0 * n_rows * DIM * B / (ms * 1e-3) / 1e12
        burst = sustained = None
        try:
            pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            burst, sustained = float(pk["bf16_tflops"]), float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
        except Exception:
            pass
        par = checker.check(lambda q: store.search_device(q, K), range(PARITY_QUERIES), batch=Qb)
        extra["c3_batch_1024"] = {
            "kernel": "gemm_filter_kernel (K2b, 128-query tcgen05 tiles) + refine_topk_kernel", "ms_per_batch": ms,
            "queries_per_s": B / ms * 1e3, "useful_tflops_all_gpus": tflops, "useful_tflops_per_gpu": tflops / world,
            "bf16_peak_burst": burst, "bf16_peak_sustained": sustained,
            "frac_of_burst": (tflops / world / burst) if burst else None,
            "frac_of_sustained": (tflops / world / sustained) if sustained else None, "parity": par}
    except Exception as e:
        extra["c3_batch_1024"] = {"error": repr(e)}
    extra["tf32_tflops_measured"] = _tf32_peak_tflops(dev)
    extra["tf32_note"] = "dense TF32 torch.matmul 8192^3, best of 6: the measured denominator for the optional 3xTF32 kernel K2"
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="debug only: a reduced matrix is flagged in config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the K1-forced / batch-1024 extra measurements")
    ap.add_argument("--cpu-sample", action="store_true", help="reference arm: scan a 1M-row sample even if RAM allows 10M")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
