set -x
timeout 900 python -m pytest tests/test_trackers_gpu.py -q -m gpu -k nccl -s > gpurun_out/r2i_nccl_test.log 2>&1; grep -A6 "first difference\|passed\|failed" gpurun_out/r2i_nccl_test.log | cut -c1-1700 | head -30
