set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel_v8 -s 1 -c 1 -f -o gpurun_out/v8_prof python scratch/prof_v8.py 8192 3 > gpurun_out/ncu_v8.log 2>&1; tail -3 gpurun_out/ncu_v8.log
