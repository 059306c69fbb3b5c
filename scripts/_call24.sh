set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_all4_b32.json 2> gpurun_out/r2z.err || tail -3 gpurun_out/r2z.err
python scripts/layer_times.py 32 > gpurun_out/r2z_layers_final.txt 2>&1; grep "==" gpurun_out/r2z_layers_final.txt
python scripts/prog_times.py 32 20 2>&1 | tail -4 > gpurun_out/r2z_prog_times.txt; cat gpurun_out/r2z_prog_times.txt
python bench.py --config pose --batch 128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_pose_b128.json 2>> gpurun_out/r2z.err
python bench.py --config ball --batch 256 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_ball_b256.json 2>> gpurun_out/r2z.err
python bench.py --res 4k --batch 64 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_4k_b64.json 2>> gpurun_out/r2z.err
python bench.py --strong --frames 4096 --steps 2 > gpurun_out/r2z_bench_strong_n1.json 2>> gpurun_out/r2z.err
PADEL_B200_NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2z_launches.csv | head -20
for f in gpurun_out/r2z_bench_*.json; do python - $f <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(sys.argv[1], d["value"], "fps", d["ms_per_step"], "ms e2e", d["e2e"]["value"], d["clocks"])
PY
done
