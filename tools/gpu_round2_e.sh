#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specialised or bitplane or both_kernel or config1 or config4 or non_ascii or edge or fused_nextseq or large_batch or random_adapter" > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
python -m pytest tests/test_gpu_gate.py -x -q -m gpu > gpurun_out/r2e_gate.log 2>&1
echo "gate rc=$?" >> gpurun_out/r2e_gate.log
tail -8 gpurun_out/r2e_gate.log
export CUTADAPT_B200_STAGE_TIMES=1
for v in jit nojit; do
  if [ $v = nojit ]; then export CUTADAPT_B200_JIT=0; else unset CUTADAPT_B200_JIT; fi
  timeout 600 python bench.py --reads 100000000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2e_bench_$v.json 2> gpurun_out/r2e_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2e_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'])"
done
unset CUTADAPT_B200_JIT
unset CUTADAPT_B200_STAGE_TIMES
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)1" -s 2 -c 1 -o gpurun_out/r2e_plan -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2e_ncu_plan.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)0" -s 8 -c 1 -o gpurun_out/r2e_run -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2e_ncu_run.log 2>&1
ls -la gpurun_out/r2e*.ncu-rep
