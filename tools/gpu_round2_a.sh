#!/bin/bash
# first GPU call of round 2: parity of the new first stage + A/B against the shift-and scan + launch lists
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bitplane or both_kernel or config1 or config4 or non_ascii or edge or fused_nextseq or large_batch" > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for v in planes shiftand; do
  if [ $v = shiftand ]; then export CUTADAPT_B200_SCAN=shiftand; else unset CUTADAPT_B200_SCAN; fi
  timeout 600 python bench.py --reads 100000000 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2a_bench_$v.json 2> gpurun_out/r2a_bench_$v.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cg_ -c 60 --csv --log-file gpurun_out/r2a_launches_$v.csv python bench.py --reads 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
done
unset CUTADAPT_B200_SCAN
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_bench_planes.json gpurun_out/r2a_bench_shiftand.json | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l); print('value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']))"
for v in planes shiftand; do echo == $v; python tools/launch_summary.py gpurun_out/r2a_launches_$v.csv | tail -12; done
