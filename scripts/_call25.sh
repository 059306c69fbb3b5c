mkdir -p gpurun_out
T="tests/test_full_size_gpu.py::test_pose_batch128_is_batch_invariant"
for env in "X=1" "PADEL_B200_CONV_EPI=0" "PADEL_B200_CONV_HALO1=0" "PADEL_B200_STEM_RAW=0" "PADEL_B200_FUSE_OUT2=0" "PADEL_B200_PIL_ROWS=1" "PADEL_B200_PDL=0"; do
  echo "== $env"; env $env timeout 600 python -m pytest $T -q -x 2>&1 | tail -2
done
