set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests.log
tail -25 gpurun_out/r2b_tests.log
timeout 300 python scripts/prog_times.py > gpurun_out/r2b_prog_default.txt 2>&1
PADEL_B200_CONV_BRES=0 timeout 300 python scripts/prog_times.py > gpurun_out/r2b_prog_bres0.txt 2>&1
PADEL_B200_CONV_DEBUG=1 timeout 300 python scripts/prog_times.py > gpurun_out/r2b_prog_plainsilu.txt 2>&1
grep "ms per" gpurun_out/r2b_prog_default.txt gpurun_out/r2b_prog_bres0.txt gpurun_out/r2b_prog_plainsilu.txt
timeout 600 python scripts/layer_times.py 32 > gpurun_out/r2b_layers.txt 2>&1
timeout 1500 python scripts/diag_tf32_floor.py 4 > gpurun_out/r2b_tf32_floor.txt 2>&1
grep -v "^\[" gpurun_out/r2b_tf32_floor.txt | tail -60
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json | cut -c1-1500
