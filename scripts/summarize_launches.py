"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
usage: summarize_launches.py launches.csv [title]"""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"<.*", "", name).replace("pb::", "")
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list summary'}")
print("# per-launch times are cold-cache / serialised under ncu; compare SHARES, not absolutes")
print(f"# total {total / 1e3:.3f} ms over {n} launches")
print("kernel,launches,total_us,share")
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{c},{us:.1f},{us / total:.4f}")
