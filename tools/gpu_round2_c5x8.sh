#!/bin/bash
# BASELINE config 5 as stated: 200 M reads sharded over 8 GPUs (25 M per GPU), final build
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c5x${N}.json 2> gpurun_out/r2c5x${N}.err
echo "rc=$?"; tail -c 900 gpurun_out/r2c5x${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c2x${N}.json 2> gpurun_out/r2c2x${N}.err
echo "rc=$?"; tail -c 700 gpurun_out/r2c2x${N}.json
