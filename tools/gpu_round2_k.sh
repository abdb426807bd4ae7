#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specialised or bitplane or both_kernel or config1 or config4 or random_adapter or statistics or large_batch or edge or non_ascii or debug" > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
export CUTADAPT_B200_STAGE_TIMES=1
for v in default; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2k_bench_$v.json 2> gpurun_out/r2k_bench_$v.err
  python -c "
import json,sys
d = json.loads(open('gpurun_out/r2k_bench_$v.json').read().strip().split('\n')[-1]); print('$v: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['config']['first_stage_specialised'])"
done
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err
python -c "
import json,sys
d = json.loads(open('gpurun_out/r2k_bench_c4.json').read().strip().split('\n')[-1]); print('config 4: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_checked'])"
