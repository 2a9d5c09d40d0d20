set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_a.json 2> gpurun_out/bench_n1_a.err; cat gpurun_out/bench_n1_a.json; tail -3 gpurun_out/bench_n1_a.err
