python scratch/time_stats2.py 2>&1 | tail -9
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; tail -3 gpurun_out/pytest_gpu4.log
timeout 600 python bench.py > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1b.json')); print(d['ms_per_step'], d['config']['main_pass_ms'], d['config']['stats_prepass_ms'], d['value'], d['e2e'])"
