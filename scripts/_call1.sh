set -x
mkdir -p gpurun_out
python -c "import torch;print(torch.cuda.get_device_name())"
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_engines_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1.log
PADEL_B200_PDL=0 timeout 300 python scripts/prog_times.py > gpurun_out/r2_prog_pdl0.txt 2>&1
PADEL_B200_PDL=1 timeout 300 python scripts/prog_times.py > gpurun_out/r2_prog_pdl1.txt 2>&1
timeout 1200 python scripts/diag_yolo_parity.py n,m 4 > gpurun_out/r2_diag_parity.txt 2>&1
for k in detect pose13; do
  timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_yolo_$k python scripts/run_yolo_once.py $k 32 > gpurun_out/r2_ncu_$k.log 2>&1
  ncu -i /tmp/r2_yolo_$k.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_yolo_$k.csv.gz
  ls -la /tmp/r2_yolo_$k.ncu-rep
done
export PADEL_B200_LIB=padel_analytics_b200/libpadel_b200_exp.so
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k experimental > gpurun_out/r2_exp_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_exp_test.log
PADEL_B200_CONV_DEBUG=8 timeout 300 python scripts/prog_times.py > gpurun_out/r2_prog_exp8.txt 2>&1
PADEL_B200_CONV_DEBUG=0 timeout 300 python scripts/prog_times.py > gpurun_out/r2_prog_exp0.txt 2>&1
tail -3 gpurun_out/r2_t1.log gpurun_out/r2_exp_test.log
cat gpurun_out/r2_prog_pdl0.txt gpurun_out/r2_prog_pdl1.txt gpurun_out/r2_prog_exp8.txt gpurun_out/r2_prog_exp0.txt
grep "^==" gpurun_out/r2_diag_parity.txt
