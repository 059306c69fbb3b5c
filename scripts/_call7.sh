set -x
mkdir -p gpurun_out
timeout 600 python bench.py --impl eager --steps 10 --warmup 3 > gpurun_out/r2g_eager.json 2> gpurun_out/r2g_eager.err; cat gpurun_out/r2g_eager.json | cut -c1-900; tail -3 gpurun_out/r2g_eager.err
timeout 900 python bench.py --strong --frames 4096 --steps 2 > gpurun_out/r2g_strong_n1.json 2> gpurun_out/r2g_strong_n1.err; cut -c1-300 gpurun_out/r2g_strong_n1.json; grep -o '"rank0_seconds": {[^}]*}' gpurun_out/r2g_strong_n1.json; tail -3 gpurun_out/r2g_strong_n1.err
timeout 900 python bench.py --config ball --batch 256 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ball256_n1.json 2> gpurun_out/r2g_ball256_n1.err; cut -c1-300 gpurun_out/r2g_ball256_n1.json; tail -3 gpurun_out/r2g_ball256_n1.err
timeout 900 python bench.py --res 4k --batch 64 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_4k64_n1.json 2> gpurun_out/r2g_4k64_n1.err; cut -c1-300 gpurun_out/r2g_4k64_n1.json; tail -3 gpurun_out/r2g_4k64_n1.err
timeout 900 python -m pytest tests/test_trackers_gpu.py tests/test_conv_gpu.py -q -m gpu -x > gpurun_out/r2g_tests.log 2>&1; tail -3 gpurun_out/r2g_tests.log
