set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-host > gpurun_out/r2e_bench_prof.json 2> gpurun_out/r2e_bench_prof.err
grep -A28 "cumulative" gpurun_out/r2e_bench_prof.err | head -40
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_bench.json").read())
print("BENCH", d["value"], d["ms_per_step"], d["timing"], d["e2e"]["value"], d["roofline"]["per_model"])
PY
PADEL_B200_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2e_launches.csv > gpurun_out/r2e_launches_summary.txt 2>&1; head -30 gpurun_out/r2e_launches_summary.txt
timeout 2000 python -m pytest tests -q -m gpu > gpurun_out/r2e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tests.log
grep -n "^E  \|^FAILED\|passed\|failed\|resnet50 court" gpurun_out/r2e_tests.log | cut -c1-300 | head -40
