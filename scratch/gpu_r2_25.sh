set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_kernel_v8 -s 1 -c 1 -f -o gpurun_out/v8_prof_r2 python scratch/prof_v8.py 32768 3 > gpurun_out/ncu_v8_r2.log 2>&1; tail -3 gpurun_out/ncu_v8_r2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:box_mean2d -s 1 -c 1 -f -o gpurun_out/box2d_prof_r2 python scratch/prof_v8.py 32768 2 > gpurun_out/ncu_box_r2.log 2>&1; tail -3 gpurun_out/ncu_box_r2.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
