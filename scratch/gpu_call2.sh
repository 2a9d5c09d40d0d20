timeout 900 python scratch/check_v6.py 16384 > gpurun_out/check_v6.log 2>&1
grep -c "^ok" gpurun_out/check_v6.log; grep "FAIL" gpurun_out/check_v6.log | head -20; tail -18 gpurun_out/check_v6.log
