set -x
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_trackers_gpu.py -q -m gpu -k nccl -s > gpurun_out/r2m_nccl_test.log 2>&1; tail -4 gpurun_out/r2m_nccl_test.log | cut -c1-300
fi
timeout 900 $TR bench.py --gpus $N --strong --frames 4096 --steps 2 > gpurun_out/r2m_strong_n$N.json 2> gpurun_out/r2m_strong_n$N.err; cut -c1-200 gpurun_out/r2m_strong_n$N.json; grep -o '"rank0_seconds": {[^}]*}' gpurun_out/r2m_strong_n$N.json; tail -2 gpurun_out/r2m_strong_n$N.err
timeout 900 $TR bench.py --gpus $N --res 4k --batch 64 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_4k64_n$N.json 2> gpurun_out/r2m_4k64_n$N.err; cut -c1-200 gpurun_out/r2m_4k64_n$N.json; tail -2 gpurun_out/r2m_4k64_n$N.err
if [ "$N" != "2" ]; then
  timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_all4_n$N.json 2> gpurun_out/r2m_all4_n$N.err; cut -c1-200 gpurun_out/r2m_all4_n$N.json; tail -2 gpurun_out/r2m_all4_n$N.err
fi
