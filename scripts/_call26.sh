mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/r2zz_bench_all4_b32.json 2> gpurun_out/r2zz.err || tail -3 gpurun_out/r2zz.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2zz_bench_reference.json 2>> gpurun_out/r2zz.err || tail -3 gpurun_out/r2zz.err
for f in gpurun_out/r2zz_bench_*.json; do python - $f <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], d["value"], d["unit"], d["ms_per_step"], "ms e2e", d["e2e"]["value"], d.get("clocks"), d.get("roofline",{}).get("frac"), d.get("cpu_baseline",{}).get("value"))
PY
done
