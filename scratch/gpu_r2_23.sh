set -x
mkdir -p gpurun_out
timeout 900 python scratch/check_v8.py --quick 16384 > gpurun_out/check_v8t.log 2>&1; grep -c "^ok" gpurun_out/check_v8t.log; grep "FAIL" gpurun_out/check_v8t.log | head; tail -4 gpurun_out/check_v8t.log
timeout 1200 python -m pytest tests/test_gpu_fused_v8.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_c.log 2>&1; tail -4 gpurun_out/pytest_c.log
python scratch/bench_window.py
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1_e.json 2> gpurun_out/bench_n1_e.err; cat gpurun_out/bench_n1_e.json
