mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for w in 0 1; do
PADEL_B200_CONV_WIDE=$w python scripts/prog_times.py 32 20 2>&1 | tail -4
PADEL_B200_CONV_WIDE=$w python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3a_bench_wide$w.json 2> gpurun_out/r3a.err || tail -3 gpurun_out/r3a.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r3a_bench_wide$w.json") if l.startswith("{")][-1])
print("wide $w :", d["value"], "fps", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
PY
done
python scripts/layer_times.py 32 > gpurun_out/r3a_layers_wide1.txt 2>&1; grep "==" gpurun_out/r3a_layers_wide1.txt
