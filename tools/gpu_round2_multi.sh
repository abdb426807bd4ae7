#!/bin/bash
# usage: gpu_round2_multi.sh N   -- multi-GPU evidence: chunk-runner test, weak and strong scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -10
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m${N}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m${N}_pytest.log; tail -4 gpurun_out/r2m${N}_pytest.log
for mode in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --scaling $mode --no-cpu > gpurun_out/r2m${N}_bench_$mode.json 2> gpurun_out/r2m${N}_bench_$mode.err
  echo "$mode rc=$?"; tail -c 1800 gpurun_out/r2m${N}_bench_$mode.json; tail -3 gpurun_out/r2m${N}_bench_$mode.err
done
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2m${N}_bench_c5.json 2> gpurun_out/r2m${N}_bench_c5.err
  echo "config 5 rc=$?"; tail -c 1500 gpurun_out/r2m${N}_bench_c5.json
fi
