set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_kernel_v6 -s 1 -c 1 -f -o gpurun_out/fused_r1c python scratch/prof_run.py 16384 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/
timeout 300 python scratch/bench_algos.py > gpurun_out/bench_algos.jsonl 2> gpurun_out/bench_algos.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
