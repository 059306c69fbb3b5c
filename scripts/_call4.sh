set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-host > gpurun_out/r2d_bench_prof.json 2> gpurun_out/r2d_bench_prof.err
grep -A45 "cumulative" gpurun_out/r2d_bench_prof.err | head -60
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2d_launches.csv > gpurun_out/r2d_launches_summary.txt 2>&1; cat gpurun_out/r2d_launches_summary.txt | head -40
timeout 900 python -m pytest tests/test_engines_gpu.py tests/test_trackers_gpu.py -q -m gpu > gpurun_out/r2d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
