set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python scratch/check_v8.py --quick 16384 > gpurun_out/check_v8s.log 2>&1; grep -c "^ok" gpurun_out/check_v8s.log; grep "FAIL" gpurun_out/check_v8s.log | head -20; tail -5 gpurun_out/check_v8s.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1_c.json 2> gpurun_out/bench_n1_c.err; cat gpurun_out/bench_n1_c.json
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_c.json 2> gpurun_out/bench_n${N}_c.err; cat gpurun_out/bench_n${N}_c.json; grep -E "step trace|pre-pass phases" gpurun_out/bench_n${N}_c.err
