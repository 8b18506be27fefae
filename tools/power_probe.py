"""Scratch: SM clock / power / throttle reasons while the batched filter runs. Usage: power_probe.py [n] [dim] [B]"""
import os, sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, wdbx_b200, pynvml
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
os.environ["WDBX_B200_GEMM_MIN_BATCH"] = "1"
eng = wdbx_b200.Engine(0, dim, "fp32", 1)
eng.reserve(0, n)
g = torch.Generator(device="cuda").manual_seed(1)
done = 0
while done < n:
    m = min(1 << 20, n - done); eng.append(0, torch.randn((m, dim), generator=g, device="cuda")); done += m
q = torch.randn((B, dim), device="cuda")
out = eng.search(q, 10, "cosine")
torch.cuda.synchronize()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)
th = threading.Thread(target=sampler); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 60
e0.record()
for _ in range(it): eng.search(q, 10, "cosine", out=out)
e1.record(); torch.cuda.synchronize()
stop = True; th.join()
ms = e0.elapsed_time(e1) / it
mid = samples[len(samples) // 4:]
clk = sorted(s[0] for s in mid); pw = sorted(s[1] for s in mid)
reasons = 0
for s in mid: reasons |= s[2]
print(f"B={B} {ms:.3f} ms {B/ms*1e3:.0f} QPS  {2.0*n*dim*B/ms/1e9:.0f} TFLOP/s | sm clock median {clk[len(clk)//2]} MHz (min {clk[0]} max {clk[-1]}), "
      f"power median {pw[len(pw)//2]:.0f} W (max {pw[-1]:.0f}), limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h)/1000:.0f} W, reasons 0x{reasons:x}, "
      f"max sm clock {pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)}")
eng.close()
