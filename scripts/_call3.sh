set -x
mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  PADEL_B200_PDL=$1 PADEL_B200_STREAMS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_pdl$1_s$2.json 2> gpurun_out/r2c_bench_pdl$1_s$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2c_bench_pdl$1_s$2.json").read())
print("PDL $1 STREAMS $2", d["value"], d["ms_per_step"], d["timing"], d["e2e"]["value"])
PY
done
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_tests.log
tail -40 gpurun_out/r2c_tests.log
timeout 1500 python scripts/diag_tf32_floor.py 4 > gpurun_out/r2c_tf32_floor.txt 2>&1
grep -v "^\[" gpurun_out/r2c_tf32_floor.txt | tail -40
