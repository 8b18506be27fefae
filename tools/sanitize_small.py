"""Tiny end-to-end run for compute-sanitizer: K4 ingest, K1 (B=1, B=8, k=100, bf16, l2), K2b, K2, K3."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))
import numpy as np, torch
import wdbx_b200

rng = np.random.default_rng(0)
for dtype, dim in (("fp32", 96), ("bf16", 200), ("fp32", 20)):
    X = rng.standard_normal((3001, dim), dtype=np.float32)
    for mode in ("0", "1"):
        os.environ["WDBX_B200_GEMM_MODE"] = mode
        os.environ["WDBX_B200_GEMM_MIN_BATCH"] = "16"
        eng = wdbx_b200.Engine(0, dim, dtype, 2)
        eng.append(0, X[:1500]); eng.append(1, X[1500:])
        eng.tombstone(0, 3); eng.overwrite(1, 7, X[0])
        for B, k, metric in ((1, 10, "cosine"), (8, 10, "l2"), (3, 100, "ip"), (40, 10, "cosine"), (130, 5, "l2")):
            Q = rng.standard_normal((B, dim), dtype=np.float32)
            s, g, c = eng.search_host(Q, k, metric=metric)
            assert c.min() == k
        qd = torch.from_numpy(rng.standard_normal((2, dim), dtype=np.float32)).cuda()
        o1 = eng.search(qd, 10); o2 = eng.search(qd, 10)
        m = eng.merge(torch.stack([o1["keys"], o2["keys"]]))
        torch.cuda.synchronize()
        eng.read_rows(0, 0, 10)
        eng.close()
print("sanitize run ok")
