#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bitplane or both_kernel or config1 or config4 or non_ascii or edge or fused_nextseq or large_batch or random_adapter" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
for v in default b7 b6; do
  if [ $v = default ]; then unset CUTADAPT_B200_LIB; else export CUTADAPT_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  timeout 600 python bench.py --reads 100000000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c_bench_$v.json 2> gpurun_out/r2c_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2c_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']))"
done
unset CUTADAPT_B200_LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cg_ -c 60 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --reads 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2c_launches.csv | tail -12
