set -x
mkdir -p gpurun_out
FSG_STEP_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8_t.json 2> gpurun_out/r2_bench_n8_t.err; cut -c1-330 gpurun_out/r2_bench_n8_t.json; grep -E "step trace|pre-pass phases" gpurun_out/r2_bench_n8_t.err | head -20
