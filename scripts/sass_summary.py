"""SASS / resource summary of every kernel in libpadel_b200.so (cuobjdump -sass, -res-usage): tcgen05 / TMEM / TMA
mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA load, UTCBAR = tcgen05.commit), MUFU counts, registers,
stack (spills), static shared memory.  Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "padel_analytics_b200/libpadel_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
dem = lambda n: re.sub(r"\(.*", "", subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()).replace("pb::", "")
usage = {dem(m.group(1)): m.groups()[1:] for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res)}
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = dem(f.split("\n", 1)[0].strip())
    c = lambda pat: len(re.findall(pat, f))
    rows.append((name, c(r"\bUTCHMMA\b(?!\.2CTA)"), c(r"UTCHMMA\.2CTA"), c(r"\bLDTM"), c(r"UTMALDG"), c(r"UTCBAR"),
                 c(r"MUFU\.EX2"), c(r"MUFU\.RCP"), c(r"STG\.\S*256")))
print(f"# SASS summary of {lib} (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)\n")
print("UTCHMMA = tcgen05.mma (cta_group::1), UTCHMMA.2CTA = cta_group::2 (CTA pairs), LDTM = tcgen05.ld (TMEM -> registers),")
print("UTMALDG = cp.async.bulk.tensor (TMA) loads, UTCBAR = tcgen05.commit.  No UTMASTG: outputs leave through 256-bit STG.\n")
print("| kernel | regs | stack B | static smem B | UTCHMMA | UTCHMMA.2CTA | LDTM | UTMALDG | UTCBAR | MUFU.EX2 | MUFU.RCP | 256-bit STG |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in sorted(rows):
    u = usage.get(r[0], ("?", "?", "?"))
    print(f"| `{r[0]}` | {u[0]} | {u[1]} | {u[2]} | " + " | ".join(str(x) for x in r[1:]) + " |")
