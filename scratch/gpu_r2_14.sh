set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -6 gpurun_out/pytest_parity.log
python scratch/bench_algos.py > gpurun_out/bench_algos_r2.jsonl 2> gpurun_out/bench_algos_r2.err; cat gpurun_out/bench_algos_r2.jsonl; tail -3 gpurun_out/bench_algos_r2.err
