set -x
mkdir -p gpurun_out
timeout 900 python scratch/check_v8.py 16384 > gpurun_out/check_v8.log 2>&1; tail -30 gpurun_out/check_v8.log
