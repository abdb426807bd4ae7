#!/bin/bash
# A/B of the chain-step pipe balance (CG_CHAIN_IMAD), fresh source-level captures of the plan and run kernels
mkdir -p gpurun_out
export CUTADAPT_B200_STAGE_TIMES=1
for v in default imad0 imad15 imad0a; do
  case $v in
    default) unset CUTADAPT_B200_JIT_DEFINES;;
    imad0) export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0";;
    imad15) export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0x15";;
    imad0a) export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0x0A";;
  esac
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2o_bench_$v.json 2> gpurun_out/r2o_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2o_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_mismatches'], d['config']['first_stage_specialised'])"
done
unset CUTADAPT_B200_JIT_DEFINES
unset CUTADAPT_B200_STAGE_TIMES
timeout 600 python -m pytest tests/test_gpu_gate.py tests/test_gpu_parity.py -q -m gpu -x -k "gate or specialised or bitplane" > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -4 gpurun_out/r2o_pytest.log
export CUTADAPT_B200_JIT=1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)1" -s 2 -c 1 -o gpurun_out/r2o_plan -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2o_ncu_plan.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)0" -s 8 -c 1 -o gpurun_out/r2o_run -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2o_ncu_run.log 2>&1
ls -la gpurun_out/r2o*.ncu-rep
