set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k stem 2>&1 | tail -4
PADEL_B200_CONV_DEBUG=8 timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k stem 2>&1 | tail -4
PADEL_B200_STEM_RAW=0 timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k stem 2>&1 | tail -2
python scripts/layer_times.py 32 pose 2>&1 | sed -n 2,4p
PADEL_B200_STEM_RAW=0 python scripts/layer_times.py 32 pose 2>&1 | sed -n 2,4p
