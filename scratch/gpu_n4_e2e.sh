set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4_numa.json 2> gpurun_out/bench_n4_numa.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4_numa.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
PY
tail -2 gpurun_out/bench_n4_numa.err
