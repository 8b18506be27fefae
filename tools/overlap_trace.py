"""Do consecutive small-batch searches overlap on the device (engine option `overlap`)?  Runs three back-to-back traced
searches on a 1.25M x 768 slice and prints, per search, when its CTAs entered and when its last CTA finished (absolute
device time, us).  Not the bench."""
import os, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "wdbx-py_b200"))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import wdbx_b200
    from wdbx_b200.engine import new_out
    n, dim = 1_250_000, 768
    eng = wdbx_b200.Engine(0, dim, "fp32", 1)
    eng.reserve(0, n)
    g = torch.Generator(device="cuda").manual_seed(1)
    done = 0
    while done < n:
        m = min(1 << 20, n - done)
        eng.append(0, torch.randn((m, dim), generator=g, device="cuda"))
        done += m
    eng.set_option("overlap", int(sys.argv[2]))
    qs = torch.randn((8, 1, dim), device="cuda")
    outs = [new_out(1, 10, qs.device) for _ in range(8)]
    for i in range(20):
        eng.search(qs[i % 8], 10, "cosine", out=outs[i % 8])
    torch.cuda.synchronize()
    os.environ["WDBX_B200_FILTER_TRACE"] = "1"
    for i in range(3):
        eng.search(qs[i], 10, "cosine", out=outs[i])
    os.environ.pop("WDBX_B200_FILTER_TRACE")
    torch.cuda.synchronize()
    sys.exit(0)

for overlap in (0, 1):
    txt = subprocess.run([sys.executable, __file__, "child", str(overlap)], capture_output=True, text=True).stdout
    rows = []
    for line in txt.splitlines():
        if line.startswith("TRACE"):
            d = dict(zip(re.findall(r"([a-z_]+) \d+", line), map(int, re.findall(r"[a-z_]+ (\d+)", line))))
            d["last"] = "LAST" in line
            rows.append(d)
    rows.sort(key=lambda r: r["entry"])
    t0 = rows[0]["entry"]
    # split into searches: a gap of more than 100 us between consecutive entries
    groups, cur = [], [rows[0]]
    for r in rows[1:]:
        if r["entry"] - cur[-1]["entry"] > 100_000:
            groups.append(cur); cur = [r]
        else:
            cur.append(r)
    groups.append(cur)
    print(f"overlap={overlap}: {len(groups)} searches")
    for gi, gset in enumerate(groups):
        ent = [r["entry"] - t0 for r in gset]
        tick = [r["entry"] - t0 + r["ticket"] for r in gset]
        last = [r for r in gset if r["last"]]
        done = (last[0]["entry"] - t0 + last[0]["done"]) / 1e3 if last else float("nan")
        print(f"  search {gi}: {len(gset)} CTAs, entry min {min(ent)/1e3:8.1f} median {sorted(ent)[len(ent)//2]/1e3:8.1f} max {max(ent)/1e3:8.1f} us;"
              f" CTA exits (ticket) min {min(tick)/1e3:8.1f} max {max(tick)/1e3:8.1f}; last CTA done {done:8.1f} us")
