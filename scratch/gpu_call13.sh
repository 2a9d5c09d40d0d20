timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "curvature or gradient or spatial or hillshade or slope" > gpurun_out/pytest_grad.log 2>&1; tail -3 gpurun_out/pytest_grad.log
timeout 300 python scratch/bench_algos.py 2>/dev/null | head -5 | cut -c1-170
