set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fused_v8.py -x -q > gpurun_out/pytest_v8.log 2>&1; tail -15 gpurun_out/pytest_v8.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; cat gpurun_out/bench_n1_b.json; tail -3 gpurun_out/bench_n1_b.err
