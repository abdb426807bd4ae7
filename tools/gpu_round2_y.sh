#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gate.py tests/test_gpu_parity.py -q -m gpu -x -k "config3 or random_adapter or both_kernel or golden or linked or wildcard" > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -4 gpurun_out/r2z_pytest.log
export CUTADAPT_B200_STAGE_TIMES=1
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2z_bench_c3.json 2> gpurun_out/r2z_bench_c3.err
python -c "
import json
d = json.loads(open('gpurun_out/r2z_bench_c3.json').read().strip().split('\n')[-1]); print('c3: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_mismatches'])"
tail -2 gpurun_out/r2z_bench_c3.err
