# the N = 1 artefacts of profiles/: bench line, reference arm, per-algorithm timings, ncu launch list of the bench command
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json | cut -c1-300
python scratch/bench_algos.py > gpurun_out/r2_bench_algos.jsonl 2> gpurun_out/r2_bench_algos.err; tail -14 gpurun_out/r2_bench_algos.jsonl | cut -c1-160
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-nodata-variant > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-120
