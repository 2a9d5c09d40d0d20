set -x
mkdir -p gpurun_out
timeout 900 python scratch/check_v8.py --quick 16384 > gpurun_out/check_v9.log 2>&1; grep -c "^ok" gpurun_out/check_v9.log; grep "FAIL" gpurun_out/check_v9.log | head -20; tail -7 gpurun_out/check_v9.log
cd fujishadergpu_b200/csrc && for f in *.cu; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr -DFSG_V8_TIMERS -I ../../include -c $f -o /tmp/$f.o & done; wait; nvcc -shared -o ../lib/libfsg_b200.so /tmp/*.cu.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart shared; cd ../..
python scratch/v8_timers.py 8192 > gpurun_out/v9_timers.log 2>&1; cat gpurun_out/v9_timers.log
