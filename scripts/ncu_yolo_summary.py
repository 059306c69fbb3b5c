"""`ncu --page raw --csv` (gzip) of one YOLO batch -> per-launch table: duration, grid, DRAM bytes, tensor pipe %,
DRAM throughput %, SM active %, achieved occupancy.  Usage: python scripts/ncu_yolo_summary.py file.csv.gz"""
import csv
import gzip
import io
import sys

rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
hdr, units, data = rows[0], rows[1], rows[2:]
c = {n: hdr.index(n) for n in hdr}
SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def g(r, n, d=0.0):
    try:
        return float(r[c[n]].replace(",", "")) * SCALE.get(units[c[n]], 1.0)
    except Exception:
        return d
tot = 0.0
print(f"{'#':>3} {'kernel':28s} {'grid':>5} {'blk':>4} {'us':>8} {'dramMB':>8} {'dram%':>6} {'tensor%':>7} {'sm_busy%':>8} {'sm_active_cyc%':>14}")
agg = {}
for i, r in enumerate(data):
    name = r[c["Kernel Name"]].split("(")[0].replace("pb::", "")[:28]
    us = g(r, "gpu__time_duration.sum") / 1e3
    mb = (g(r, "dram__bytes_read.sum") + g(r, "dram__bytes_write.sum")) / 1e6
    el = g(r, "sm__cycles_elapsed.max")
    act = g(r, "smsp__cycles_active.avg")
    print(f"{i:3d} {name:28s} {int(g(r,'launch__grid_size')):5d} {int(g(r,'launch__block_size')):4d} {us:8.1f} {mb:8.2f} "
          f"{g(r,'dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):7.1f} "
          f"{g(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed'):8.1f} {100*act/el if el else 0:14.1f}")
    tot += us
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] += mb
print(f"total {tot:.1f} us over {len(data)} launches (serialised, cold-cache ncu replays)")
for k, (n, us, mb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} x{n:3d} {us:9.1f} us {mb:9.1f} MB  {100*us/tot:5.1f}%")
