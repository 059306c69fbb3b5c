set -x
mkdir -p gpurun_out
python scripts/layer_times.py 32 pose,players,court > gpurun_out/r2x_layers_raw.txt 2>&1
PADEL_B200_STEM_RAW=0 python scripts/layer_times.py 32 pose,players,court > gpurun_out/r2x_layers_box.txt 2>&1
grep -A1 "==" gpurun_out/r2x_layers_raw.txt gpurun_out/r2x_layers_box.txt
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_engines_gpu.py -q -m gpu -x 2>&1 | tail -3
python scripts/prog_times.py 32 20 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x.err || tail -3 gpurun_out/r2x.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2x_bench.json") if l.startswith("{")][-1])
print("bench :", d["value"], "fps", d["ms_per_step"], "ms  e2e", d["e2e"]["value"])
PY
