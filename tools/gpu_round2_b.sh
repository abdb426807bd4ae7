#!/bin/bash
# ncu source-level profile of cg_pscan_kernel + timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bitplane or both_kernel" > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 600 python bench.py --reads 100000000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cg_ -c 60 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --reads 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:cg_pscan -s 2 -c 1 -o gpurun_out/r2b_pscan -f python bench.py --reads 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
tail -3 gpurun_out/r2b_pytest.log
python -c "
import json
d = json.loads(open('gpurun_out/r2b_bench.json').read().strip().split('\n')[-1]); print('value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']))"
python tools/launch_summary.py gpurun_out/r2b_launches.csv | tail -12
ls -la gpurun_out/r2b_pscan.ncu-rep
