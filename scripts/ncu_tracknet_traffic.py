"""Turn an `ncu --set full` capture of scripts/run_tracknet_once.py into per-launch DRAM traffic + tensor-pipe numbers.

  ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_tracknet_traffic.py raw.csv B
"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2])
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
def col(r, name):
    try:
        return float(r[idx[name]].replace(",", ""))
    except Exception:
        return float("nan")
out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    out.append(dict(kernel=r[idx["Kernel Name"]].split("(")[0], us=col(r, "gpu__time_duration.sum") / 1e3 if col(r, "gpu__time_duration.sum") > 1e4 else col(r, "gpu__time_duration.sum"),
                    dram_read_mb=col(r, "dram__bytes_read.sum"), dram_write_mb=col(r, "dram__bytes_write.sum"),
                    tensor_pct=col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    dram_pct=col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    l2_hit=col(r, "lts__t_sector_hit_rate.pct"), regs=col(r, "launch__registers_per_thread")))
print(json.dumps(out, indent=1))
