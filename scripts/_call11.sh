set -x
mkdir -p gpurun_out
PADEL_B200_CONV_OCC2=2 timeout 300 python scripts/exp_chain_timeline.py > gpurun_out/r2k_chain_occ2.txt 2>&1; cat gpurun_out/r2k_chain_occ2.txt | grep -v "predictions"
PADEL_B200_CONV_OCC2=2 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_engines_gpu.py -q -m gpu -x > gpurun_out/r2k_tests_occ2.log 2>&1; tail -3 gpurun_out/r2k_tests_occ2.log
for m in 1 2; do PADEL_B200_CONV_OCC2=$m timeout 300 python scripts/prog_times.py > gpurun_out/r2k_prog_occ$m.txt 2>&1; grep "ms per" gpurun_out/r2k_prog_occ$m.txt; done
PADEL_B200_CONV_OCC2=2 PADEL_B200_PDL=0 timeout 300 python scripts/prog_times.py > gpurun_out/r2k_prog_occ2_pdl0.txt 2>&1; grep "ms per" gpurun_out/r2k_prog_occ2_pdl0.txt
PADEL_B200_CONV_OCC2=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench_occ2.json 2> gpurun_out/r2k_bench_occ2.err; cut -c1-250 gpurun_out/r2k_bench_occ2.json
