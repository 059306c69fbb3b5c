set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_engines_gpu.py -q -m gpu -x 2>&1 | tail -4
python scripts/layer_times.py 32 pose > gpurun_out/r2y_layers_h1.txt 2>&1
PADEL_B200_CONV_HALO1=0 python scripts/layer_times.py 32 pose > gpurun_out/r2y_layers_h0.txt 2>&1
grep "==" gpurun_out/r2y_layers_h?.txt
for h in 0 1; do
PADEL_B200_CONV_HALO1=$h python scripts/prog_times.py 32 20 2>&1 | tail -4
PADEL_B200_CONV_HALO1=$h python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_bench_h$h.json 2> gpurun_out/r2y.err || tail -3 gpurun_out/r2y.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2y_bench_h$h.json") if l.startswith("{")][-1])
print("halo1x1 $h :", d["value"], "fps", d["ms_per_step"], "ms  e2e", d["e2e"]["value"])
PY
done
