set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m pytest tests/test_gpu_sharding.py -x -q > gpurun_out/pytest_shard.log 2>&1; tail -8 gpurun_out/pytest_shard.log
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_d.json 2> gpurun_out/bench_n${N}_d.err; cat gpurun_out/bench_n${N}_d.json; grep -E "step trace|pre-pass phases|Warn|warn" gpurun_out/bench_n${N}_d.err | head -20
FSG_NO_PEER=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_e.json 2> gpurun_out/bench_n${N}_e.err; cat gpurun_out/bench_n${N}_e.json | cut -c1-400
