"""ncu driver: B queries per call on an n x dim fp32 matrix. Usage: profile_batch.py B [n] [dim] [launches]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import torch, wdbx_b200
B = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 768; launches = int(sys.argv[4]) if len(sys.argv) > 4 else 4
eng = wdbx_b200.Engine(0, dim, "fp32", 1)
g = torch.Generator(device="cuda").manual_seed(1)
done = 0
while done < n:
    m = min(1 << 20, n - done); eng.append(0, torch.randn((m, dim), generator=g, device="cuda")); done += m
q = torch.randn((B, dim), device="cuda")
out = eng.search(q, 10, "cosine")
for _ in range(launches): eng.search(q, 10, "cosine", out=out)
torch.cuda.synchronize()
print("ok", out["gids"][0, :3].tolist())
eng.close()
