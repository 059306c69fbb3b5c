set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 900 python -m pytest tests/test_trackers_gpu.py -q -m gpu -k nccl -s > gpurun_out/r2h_nccl_test.log 2>&1; tail -5 gpurun_out/r2h_nccl_test.log
timeout 900 $TR bench.py --gpus 2 --config ball --batch 256 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_ball256_n2.json 2> gpurun_out/r2h_ball256_n2.err; cut -c1-300 gpurun_out/r2h_ball256_n2.json; tail -2 gpurun_out/r2h_ball256_n2.err
timeout 900 $TR bench.py --gpus 2 --res 4k --batch 64 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_4k64_n2.json 2> gpurun_out/r2h_4k64_n2.err; cut -c1-300 gpurun_out/r2h_4k64_n2.json; tail -2 gpurun_out/r2h_4k64_n2.err
timeout 900 $TR bench.py --gpus 2 --strong --frames 4096 --steps 2 > gpurun_out/r2h_strong_n2.json 2> gpurun_out/r2h_strong_n2.err; cut -c1-300 gpurun_out/r2h_strong_n2.json; grep -o '"rank0_seconds": {[^}]*}' gpurun_out/r2h_strong_n2.json; tail -2 gpurun_out/r2h_strong_n2.err
