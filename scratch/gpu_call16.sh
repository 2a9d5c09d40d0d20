timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; tail -3 gpurun_out/pytest_gpu6.log
timeout 900 python bench.py > gpurun_out/final4_bench_n1.json 2> gpurun_out/final4_bench_n1.err; tail -2 gpurun_out/final4_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final4_bench_n1.json') if l.startswith('{')][0])
print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), 'main', round(d['config']['main_pass_ms'],2), 'stats(step-main)', round(d['config']['stats_prepass_ms'],2), 'fused', round(d['config']['fused_kernel_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],1), d['config']['scale_p99'], d['gpu_launches'], d['roofline']['frac'])
PY
