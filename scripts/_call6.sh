set -x
mkdir -p gpurun_out
timeout 300 python scripts/layer_times.py 32 ball > gpurun_out/r2f_layers_ball_default.txt 2>&1
PADEL_B200_CONV_PAIR=1 timeout 300 python scripts/layer_times.py 32 ball > gpurun_out/r2f_layers_ball_pair1.txt 2>&1
PADEL_B200_CONV_BRES=0 timeout 300 python scripts/layer_times.py 32 ball > gpurun_out/r2f_layers_ball_bres0.txt 2>&1
paste <(grep conv gpurun_out/r2f_layers_ball_default.txt | awk '{print $1,$3,$10,$11,$12,$13}') <(grep conv gpurun_out/r2f_layers_ball_pair1.txt | awk '{print $3}') <(grep conv gpurun_out/r2f_layers_ball_bres0.txt | awk '{print $3}')
timeout 600 python bench.py --impl eager --steps 10 --warmup 3 > gpurun_out/r2f_eager.json 2> gpurun_out/r2f_eager.err; cat gpurun_out/r2f_eager.json | cut -c1-900; tail -3 gpurun_out/r2f_eager.err
timeout 900 python bench.py --config pose --batch 128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_pose128.json 2> gpurun_out/r2f_bench_pose128.err; cut -c1-400 gpurun_out/r2f_bench_pose128.json; tail -3 gpurun_out/r2f_bench_pose128.err
timeout 900 python bench.py --res 4k --batch 32 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_4k.json 2> gpurun_out/r2f_bench_4k.err; cut -c1-400 gpurun_out/r2f_bench_4k.json; tail -3 gpurun_out/r2f_bench_4k.err
timeout 900 python bench.py --strong --frames 1024 --steps 2 > gpurun_out/r2f_bench_strong1.json 2> gpurun_out/r2f_bench_strong1.err; cut -c1-1200 gpurun_out/r2f_bench_strong1.json; tail -5 gpurun_out/r2f_bench_strong1.err
