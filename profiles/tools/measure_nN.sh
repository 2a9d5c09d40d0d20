set -x
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cat gpurun_out/r2_bench_n${N}.json | cut -c1-330; grep -E "Warn|warn|Error" gpurun_out/r2_bench_n${N}.err | head -5
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_sharding.py -x -q > gpurun_out/pytest_shard2.log 2>&1; tail -3 gpurun_out/pytest_shard2.log; fi
