set -x
mkdir -p gpurun_out
N=${1:-8}
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_k.json 2> gpurun_out/bench_n${N}_k.err; cat gpurun_out/bench_n${N}_k.json; grep -E "step trace|pre-pass phases|Warn|warn|Error" gpurun_out/bench_n${N}_k.err | head -20
