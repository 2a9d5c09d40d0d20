set -x
mkdir -p gpurun_out
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scratch/symm_probe.py 2>&1 | grep -E "copy|NCCL s"
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_h.json 2> gpurun_out/bench_n${N}_h.err; cut -c1-330 gpurun_out/bench_n${N}_h.json; grep -E "step trace|pre-pass phases|Warn|warn" gpurun_out/bench_n${N}_h.err | head -20
