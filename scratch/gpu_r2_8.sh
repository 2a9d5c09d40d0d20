set -x
mkdir -p gpurun_out
N=${1:-2}
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_a.json 2> gpurun_out/bench_n${N}_a.err; cat gpurun_out/bench_n${N}_a.json; grep -E "step trace|pre-pass phases" gpurun_out/bench_n${N}_a.err
if [ "$N" = "2" ]; then timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m pytest tests/test_gpu_sharding.py -x -q > gpurun_out/pytest_shard2.log 2>&1; tail -5 gpurun_out/pytest_shard2.log; fi
