set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "pil or natural" 2>&1 | tail -4
for r in 1 4; do
PADEL_B200_PIL_ROWS=$r python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_rows$r.json 2> gpurun_out/r2t.err || tail -3 gpurun_out/r2t.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2t_bench_rows$r.json") if l.startswith("{")][-1])
print("pil_rows $r :", d["value"], "fps", d["ms_per_step"], "ms  e2e", d["e2e"]["value"])
PY
done
PADEL_B200_NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2t_launches.csv | head -24
