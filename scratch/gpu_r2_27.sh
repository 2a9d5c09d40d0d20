set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_fused_v8.py -x -q > gpurun_out/pytest_e.log 2>&1; tail -5 gpurun_out/pytest_e.log
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_l.json 2> gpurun_out/bench_n${N}_l.err; cut -c1-330 gpurun_out/bench_n${N}_l.json; grep -E "step trace|pre-pass phases|Warn|warn|Error" gpurun_out/bench_n${N}_l.err | head -20
