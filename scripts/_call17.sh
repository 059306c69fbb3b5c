set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -o gpurun_out/r2r_halo16 -f python scripts/run_conv_once.py 32 320 320 16 16 3 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/r2r_tc32 -f python scripts/run_conv_once.py 32 320 320 32 32 1 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
