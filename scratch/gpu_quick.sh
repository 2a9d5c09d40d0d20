set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_v8.py tests/test_gpu_fullsize.py -x -q -k "percentile or stats or scale or sync_free or nine or fullsize or p99" > gpurun_out/pytest_pct.log 2>&1; tail -4 gpurun_out/pytest_pct.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-nodata-variant > gpurun_out/bench_n1_h.json 2> gpurun_out/bench_n1_h.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_h.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['config']['main_pass_ms'], d['config']['stats_prepass_ms'], d['config']['kernel_ms_per_step'], d['config']['out_checksum'], d['config']['scale_p99'])
PY
tail -3 gpurun_out/bench_n1_h.err
