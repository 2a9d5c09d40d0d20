set -x
mkdir -p gpurun_out
timeout 900 python scratch/check_v8.py --quick 16384 > gpurun_out/check_v8.log 2>&1; grep -c "^ok" gpurun_out/check_v8.log; grep "FAIL" gpurun_out/check_v8.log | head -20; tail -5 gpurun_out/check_v8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel_v8 -s 1 -c 1 -f -o gpurun_out/v8_prof python scratch/prof_v8.py 8192 3 > gpurun_out/ncu_v8.log 2>&1; tail -3 gpurun_out/ncu_v8.log
