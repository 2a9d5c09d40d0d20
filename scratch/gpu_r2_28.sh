set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; tail -6 gpurun_out/pytest_gpu_full.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; cat gpurun_out/bench_n1_final.json; tail -3 gpurun_out/bench_n1_final.err
