timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -2 gpurun_out/bench_n4.err; cat gpurun_out/bench_n4.json | cut -c1-900
