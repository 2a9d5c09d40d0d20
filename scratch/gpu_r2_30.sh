set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full3.log 2>&1; tail -3 gpurun_out/pytest_gpu_full3.log
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nodata2.csv python scratch/prof_nodata.py > gpurun_out/prof_nodata.log 2>&1; tail -2 gpurun_out/prof_nodata.log
