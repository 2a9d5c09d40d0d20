N=$1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; cat gpurun_out/final_bench_n1.json
  timeout 600 python scratch/bench_algos.py > gpurun_out/final_bench_algos.jsonl 2>/dev/null; cat gpurun_out/final_bench_algos.jsonl | cut -c1-200
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; cat gpurun_out/final_bench_n$N.json
fi
