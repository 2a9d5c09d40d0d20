set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:openness_interior -s 1 -c 1 -f -o gpurun_out/open_prof python scratch/prof_open.py 8192 > gpurun_out/ncu_open.log 2>&1; tail -2 gpurun_out/ncu_open.log
