"""Minimal driver for ncu: fill one engine with the C3 (10M x 768 fp32) or C2 (1M x 384) matrix and
launch the scan kernel a few times.  Usage: python tools/profile_scan.py [c2|c3|c4] [launches]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch
import wdbx_b200

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n, dim, dt, metric, k = {"c2": (1_000_000, 384, "fp32", "cosine", 10), "c3": (10_000_000, 768, "fp32", "cosine", 10),
                         "c4": (12_500_000, 384, "bf16", "ip", 100)}[cfg]
eng = wdbx_b200.Engine(0, dim, dt, 1)
eng.reserve(0, n)
g = torch.Generator(device="cuda").manual_seed(1)
done = 0
while done < n:
    m = min(1 << 20, n - done)
    eng.append(0, torch.randn((m, dim), generator=g, device="cuda"))
    done += m
q = torch.randn((8, dim), device="cuda")
out = eng.search(q[0:1], k, metric)
for i in range(launches):
    eng.search(q[i % 8:i % 8 + 1], k, metric, out=out)
torch.cuda.synchronize()
print(cfg, "ok", out["gids"][0, :3].tolist(), eng.stats()["kernel_launches"])
eng.close()
