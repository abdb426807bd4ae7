#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gate.py tests/test_gpu_parity.py -q -m gpu -x -k "gate or specialised or bitplane or config4 or random_adapter or both_kernel" > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log; tail -5 gpurun_out/r2t_pytest.log
export CUTADAPT_B200_STAGE_TIMES=1
show() {
python -c "
import json,sys
d = json.loads(open('gpurun_out/r2t_bench_$1.json').read().strip().split('\n')[-1]); print('$1: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_mismatches'], d['config']['first_stage_specialised'])"
}
for v in default noband default2; do
  unset CUTADAPT_B200_LIB
  case $v in
    noband) export CUTADAPT_B200_LIB=$PWD/build_variants/lib_noband.so;;
  esac
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2t_bench_$v.json 2> gpurun_out/r2t_bench_$v.err
  show $v
done
for v in c4 c4_noband c4_b2; do
  unset CUTADAPT_B200_LIB
  case $v in
    c4_noband) export CUTADAPT_B200_LIB=$PWD/build_variants/lib_noband.so;;
    c4_b2) export CUTADAPT_B200_LIB=$PWD/build_variants/lib_run48b2.so;;
  esac
  timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2t_bench_$v.json 2> gpurun_out/r2t_bench_$v.err
  show $v
done
