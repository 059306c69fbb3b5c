set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "secondary or pair" 2>&1 | tail -5
timeout 1500 python -m pytest tests/test_engines_gpu.py tests/test_trackers_gpu.py -q -m gpu -x 2>&1 | tail -5
PADEL_B200_FUSE_OUT2=0 python scripts/prog_times.py 32 20 2>&1 | tail -5 > gpurun_out/r2n_prog_nofuse.txt; cat gpurun_out/r2n_prog_nofuse.txt
python scripts/prog_times.py 32 20 2>&1 | tail -5 > gpurun_out/r2n_prog_fuse.txt; cat gpurun_out/r2n_prog_fuse.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; cut -c1-400 gpurun_out/r2n_bench.json; tail -2 gpurun_out/r2n_bench.err
