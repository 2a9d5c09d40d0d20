set -x
mkdir -p gpurun_out
python scratch/bench_grad.py > gpurun_out/bench_grad.log 2>&1; cat gpurun_out/bench_grad.log
timeout 600 ncu --set full --clock-control none -k regex:grad_stream_kernel -s 2 -c 1 -f -o gpurun_out/grad_prof python scratch/bench_grad.py > gpurun_out/ncu_grad.log 2>&1; tail -2 gpurun_out/ncu_grad.log
