set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused_v8.py -x -q > gpurun_out/pytest_v8b.log 2>&1; tail -15 gpurun_out/pytest_v8b.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity_b.log 2>&1; tail -5 gpurun_out/pytest_parity_b.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; cat gpurun_out/bench_n1_d.json
