"""Phase timeline of ONE small-batch filter launch (gemm_filter_small_kernel): every CTA prints its own
timestamps (WDBX_B200_FILTER_TRACE=1); this tool runs one traced search on a 1.25M x 768 slice (what one rank
holds at N = 8) and summarises where the launch's microseconds go.  Not the bench."""
import os, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import wdbx_b200
    n, dim = int(sys.argv[2]), int(sys.argv[3])
    eng = wdbx_b200.Engine(0, dim, "fp32", 1)
    eng.reserve(0, n)
    g = torch.Generator(device="cuda").manual_seed(1)
    done = 0
    while done < n:
        m = min(1 << 20, n - done)
        eng.append(0, torch.randn((m, dim), generator=g, device="cuda"))
        done += m
    qs = torch.randn((8, 1, dim), device="cuda")
    out = eng.search(qs[0], 10, "cosine")
    for i in range(10):
        eng.search(qs[i % 8], 10, "cosine", out=out)
    torch.cuda.synchronize()
    os.environ["WDBX_B200_FILTER_TRACE"] = "1"
    eng.search(qs[3], 10, "cosine", out=out)
    torch.cuda.synchronize()
    os.environ.pop("WDBX_B200_FILTER_TRACE")
    sys.exit(0)

n, dim = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1_250_000, 768)
txt = subprocess.run([sys.executable, __file__, "child", str(n), str(dim)], capture_output=True, text=True).stdout
rows = []
for line in txt.splitlines():
    if line.startswith("TRACE"):
        d = dict(zip(re.findall(r"([a-z_]+) \d+", line), map(int, re.findall(r"[a-z_]+ (\d+)", line))))
        d["last"] = "LAST" in line
        rows.append(d)
per_cta = {}
for line in txt.splitlines():
    if line.startswith("TILES"):
        parts = line.split()
        cta, i0 = int(parts[2]), int(parts[4])
        per_cta.setdefault(cta, {}).update({i0 + j: int(v) for j, v in enumerate(parts[6:14])})
if per_cta:
    print("time at which the accumulator of the CTA's n-th tile was ready (us after the CTA's entry), and the step between tiles:")
    for cta in sorted(per_cta)[:4]:
        ts = [per_cta[cta][i] for i in sorted(per_cta[cta]) if per_cta[cta][i] > 0]
        print(f"  cta {cta:3d}: first {ts[0]/1e3:.1f}  " + " ".join(f"{(b - a)/1e3:.1f}" for a, b in zip(ts, ts[1:])))
if not rows:
    print(txt[-2000:])
    sys.exit(1)
t0 = min(r["entry"] for r in rows)
def col(name, fn=lambda r, v: v):
    v = sorted(r["entry"] - t0 + r[name] for r in rows if name in r)
    return f"{name:>9}: min {v[0]/1e3:7.1f}  median {v[len(v)//2]/1e3:7.1f}  max {v[-1]/1e3:7.1f} us"
print(f"{len(rows)} CTAs, {n} x {dim}; times since the first CTA's entry")
ent = sorted(r["entry"] - t0 for r in rows)
print(f"    entry: min {ent[0]/1e3:7.1f}  median {ent[len(ent)//2]/1e3:7.1f}  max {ent[-1]/1e3:7.1f} us")
for name in ("loop_end", "tail_in", "q_ready", "rescored", "fenced", "ticket"):
    print(col(name))
last = [r for r in rows if r["last"]][0]
print(f"last CTA {last['cta']}: " + "  ".join(f"{n} {(last['entry'] - t0 + last[n]) / 1e3:.1f}" for n in
      ("loop_end", "tail_in", "q_ready", "rescored", "fenced", "ticket", "flags", "merged", "done")) + f" us; final keys {last['keys']}")
c = sorted(r["cand"] for r in rows)
print(f"candidates per CTA: min {c[0]} median {c[len(c)//2]} max {c[-1]} sum {sum(c)}")
