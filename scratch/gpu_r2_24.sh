set -x
mkdir -p gpurun_out
timeout 900 python scratch/check_v8.py --quick 16384 > gpurun_out/check_v8u.log 2>&1; grep -c "^ok" gpurun_out/check_v8u.log; grep "FAIL" gpurun_out/check_v8u.log | head; tail -4 gpurun_out/check_v8u.log
timeout 1200 python -m pytest tests/test_gpu_fused_v8.py -x -q > gpurun_out/pytest_d.log 2>&1; tail -4 gpurun_out/pytest_d.log
python scratch/bench_window.py
