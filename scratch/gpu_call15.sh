timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/pytest_shard2.log 2>&1; tail -3 gpurun_out/pytest_shard2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final3_bench_n2.json 2> gpurun_out/final3_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final3_bench_n2.json') if l.startswith('{')][0])
print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), 'main', round(d['config']['main_pass_ms'],2), 'stats(step-main)', round(d['config']['stats_prepass_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],1), d['config']['scale_p99'])
PY
tail -3 gpurun_out/final3_bench_n2.err
