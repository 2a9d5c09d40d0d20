set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full2.log 2>&1; tail -4 gpurun_out/pytest_gpu_full2.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1_f.json 2> gpurun_out/bench_n1_f.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_f.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['config']['nodata_variant'], d['config']['kernel_ms_per_step'])
PY
