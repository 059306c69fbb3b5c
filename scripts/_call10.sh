set -x
mkdir -p gpurun_out
timeout 300 python scripts/exp_chain_timeline.py > gpurun_out/r2j_chain_pdl1.txt 2>&1; cat gpurun_out/r2j_chain_pdl1.txt | grep -v "predictions"
PADEL_B200_PDL=0 timeout 300 python scripts/exp_chain_timeline.py > gpurun_out/r2j_chain_pdl0.txt 2>&1; grep -A9 "P4 bottleneck" gpurun_out/r2j_chain_pdl0.txt
timeout 900 python bench.py --strong --frames 4096 --steps 2 > gpurun_out/r2j_strong_n1.json 2> gpurun_out/r2j_strong_n1.err; cut -c1-200 gpurun_out/r2j_strong_n1.json; grep -o '"rank0_seconds": {[^}]*}' gpurun_out/r2j_strong_n1.json; tail -2 gpurun_out/r2j_strong_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; cut -c1-250 gpurun_out/r2j_bench.json; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/r2j_bench.json
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2j_tests.log 2>&1; tail -4 gpurun_out/r2j_tests.log
