#!/bin/bash
mkdir -p gpurun_out
export CUTADAPT_B200_STAGE_TIMES=1
for v in default pb6; do
  if [ $v = default ]; then unset CUTADAPT_B200_LIB; else export CUTADAPT_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  timeout 600 python bench.py --reads 100000000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2g_bench_$v.json 2> gpurun_out/r2g_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2g_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_checked'])"
done
unset CUTADAPT_B200_LIB
for c in 3 4 5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2g_bench_c$c.json 2> gpurun_out/r2g_bench_c$c.err
  tail -2 gpurun_out/r2g_bench_c$c.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2g_bench_c$c.json').read().strip().split('\n')[-1]); print('config $c: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_checked'], d['config']['with_adapters'])"
done
