#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "statistics or specialised or config5 or comparers" > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -5 gpurun_out/r2q_pytest.log
export CUTADAPT_B200_STAGE_TIMES=1
show() {
python -c "
import json,sys
d = json.loads(open('gpurun_out/r2q_bench_$1.json').read().strip().split('\n')[-1]); print('$1: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_mismatches'], d['config']['first_stage_specialised'])"
}
for v in default nofuse nofuse_imad0 imad0 default2; do
  unset CUTADAPT_B200_JIT_DEFINES CUTADAPT_B200_NO_FUSED_STATS
  case $v in
    imad0) export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0";;
    nofuse) export CUTADAPT_B200_NO_FUSED_STATS=1;;
    nofuse_imad0) export CUTADAPT_B200_NO_FUSED_STATS=1; export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0";;
  esac
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2q_bench_$v.json 2> gpurun_out/r2q_bench_$v.err
  show $v
done
