#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gate.py tests/test_gpu_parity.py tests/test_gpu_fastq.py -q -m gpu -x -k "config5 or index or demultiplex or barcode" > gpurun_out/r2last3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2last3_pytest.log; tail -4 gpurun_out/r2last3_pytest.log
timeout 200 python bench.py --config 5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2last3_c5.json 2>/dev/null
python -c "
import json
d = json.loads(open('gpurun_out/r2last3_c5.json').read().strip().split('\n')[-1]); print('c5: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['parity_mismatches'], d['parity_checked'])"
