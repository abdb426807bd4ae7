#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "debug_matrices" 2>&1 | tail -3
export CUTADAPT_B200_STAGE_TIMES=1
for v in default onelist nojit; do
  unset CUTADAPT_B200_ONE_LIST CUTADAPT_B200_JIT
  if [ $v = onelist ]; then export CUTADAPT_B200_ONE_LIST=1; fi
  if [ $v = nojit ]; then export CUTADAPT_B200_JIT=0; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2j_bench_$v.json 2> gpurun_out/r2j_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2j_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['config']['first_stage_specialised'], d['clocks'])"
done
