set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python bench.py --strong --frames 4096 --steps 2 > gpurun_out/r2o_strong_n1.json 2> gpurun_out/r2o_strong_n1.err; cut -c1-200 gpurun_out/r2o_strong_n1.json; grep -o '"rank0_seconds": {[^}]*}' gpurun_out/r2o_strong_n1.json; tail -2 gpurun_out/r2o_strong_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; cat gpurun_out/r2o_bench.json; tail -2 gpurun_out/r2o_bench.err
