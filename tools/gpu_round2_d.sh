#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specialised or bitplane or both_kernel or config1 or config4 or non_ascii or edge or fused_nextseq or large_batch or random_adapter" > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
for v in jit nojit; do
  if [ $v = nojit ]; then export CUTADAPT_B200_JIT=0; else unset CUTADAPT_B200_JIT; fi
  timeout 600 python bench.py --reads 100000000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2d_bench_$v.json 2> gpurun_out/r2d_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2d_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']))"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cg_ -c 80 --csv --log-file gpurun_out/r2d_launches_$v.csv python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r2d_launches_$v.csv | tail -9
done
unset CUTADAPT_B200_JIT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:cg_list_kernel -s 7 -c 1 -o gpurun_out/r2d_plan -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:cg_pscan -s 2 -c 1 -o gpurun_out/r2d_pscan_jit -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
