set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cog_io.py -x -q -k "sharded" > gpurun_out/pytest_cogsh.log 2>&1; tail -5 gpurun_out/pytest_cogsh.log
python profiles/tools/run_config5.py --size 16384 --out /tmp/c5_n1.tif > gpurun_out/c5_16k_n1.json 2> gpurun_out/c5_16k_n1.err; cat gpurun_out/c5_16k_n1.json; tail -3 gpurun_out/c5_16k_n1.err
