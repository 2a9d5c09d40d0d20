set -x
mkdir -p gpurun_out
N=${1:-4}
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_f.json 2> gpurun_out/bench_n${N}_f.err; cut -c1-330 gpurun_out/bench_n${N}_f.json; grep -E "step trace|pre-pass phases|Warn|warn" gpurun_out/bench_n${N}_f.err | head -20
FSG_STEP_TRACE=1 FSG_NO_PEER=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_g.json 2> gpurun_out/bench_n${N}_g.err; cut -c1-330 gpurun_out/bench_n${N}_g.json ; grep -E "step trace|pre-pass phases|Warn|warn" gpurun_out/bench_n${N}_g.err | head -20
