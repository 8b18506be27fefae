"""ncu driver for K2. Usage: profile_gemm.py [n] [dim] [B]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, wdbx_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
os.environ["WDBX_B200_GEMM_MIN_BATCH"] = "1"
eng = wdbx_b200.Engine(0, dim, "fp32", 1)
g = torch.Generator(device="cuda").manual_seed(1)
done = 0
while done < n:
    m = min(1 << 20, n - done); eng.append(0, torch.randn((m, dim), generator=g, device="cuda")); done += m
q = torch.randn((B, dim), device="cuda")
out = eng.search(q, 10, "cosine")
for _ in range(3): eng.search(q, 10, "cosine", out=out)
torch.cuda.synchronize()
print("ok", out["gids"][0, :3].tolist())
eng.close()
