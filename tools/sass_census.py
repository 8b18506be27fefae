"""profiles/sass_opcodes.txt: per-kernel census of the Blackwell-native SASS opcodes in libwdbx_b200.so
(cuobjdump -sass; runs without a GPU).  Evidence that the hot kernels are tcgen05 / TMEM / TMA code."""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "wdbx-py_b200" / "wdbx_b200" / "libwdbx_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[.0-9A-Z]*|LDTM[.0-9A-Zx]*|UTMALDG[.0-9A-Z]*|UBLKCP[.A-Z]*|UTCBAR[.0-9A-Z]*|SYNCS[.A-Z0-9]*"
                 r"|UTCATOMSWS[.A-Z0-9]*|ACQBULK|PREEXIT|ELECT|UTMAPF[.A-Z0-9]*|UTMACCTL[.A-Z0-9]*)\b")
out = ["# SASS opcode census of libwdbx_b200.so (cuobjdump -sass, sm_100a), per kernel; tools/sass_census.py.",
       "# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), other UTC*MMA = other tcgen05.mma kinds (kind::tf32 = UTCHMMA too),",
       "# LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (2-D TMA), UBLKCP = cp.async.bulk (1-D TMA),",
       "# UTCBAR = tcgen05.commit, SYNCS.* = mbarrier operations, UTCATOMSWS = tcgen05.alloc / dealloc,",
       "# ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch), ELECT = elect.sync.", ""]
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"wdbx::\(anonymous namespace\)::", "", dem)
    dem = re.sub(r"\(.*$", "", dem)
    c = collections.Counter(m.group(1) for m in pat.finditer(f))
    if c:
        out.append(dem)
        out += [f"    {v:6d}  {k}" for k, v in sorted(c.items())]
(ROOT / "profiles" / "sass_opcodes.txt").write_text("\n".join(out) + "\n")
print(f"{len(out)} lines -> profiles/sass_opcodes.txt")
