set -x
mkdir -p gpurun_out
N=${1:-8}
SIZE=${2:-131072}
EXTRA=${3:-}
if [ "$N" = "1" ]; then
python profiles/tools/run_config5.py --size $SIZE --out /tmp/c5.tif $EXTRA > gpurun_out/c5_${SIZE}_n${N}.json 2> gpurun_out/c5_${SIZE}_n${N}.err
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 profiles/tools/run_config5.py --gpus $N --size $SIZE --out /tmp/c5.tif $EXTRA > gpurun_out/c5_${SIZE}_n${N}.json 2> gpurun_out/c5_${SIZE}_n${N}.err
fi
cat gpurun_out/c5_${SIZE}_n${N}.json; tail -3 gpurun_out/c5_${SIZE}_n${N}.err
