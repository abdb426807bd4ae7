#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gate.py tests/test_gpu_parity.py tests/test_gpu_fastq.py -q -m gpu -x -k "config5 or index or demultiplex or barcode or non_ascii or edge or large_batch" > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log; tail -5 gpurun_out/r2v_pytest.log
show() {
python -c "
import json,sys
d = json.loads(open('gpurun_out/r2v_bench_$1.json').read().strip().split('\n')[-1]); print('$1: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['parity_mismatches'], d['parity_checked'])"
}
for v in c5 c5_noindex c5_nolight; do
  unset CUTADAPT_B200_NO_LIGHT CUTADAPT_B200_NO_INDEX_KERNEL
  if [ $v = c5_nolight ]; then export CUTADAPT_B200_NO_LIGHT=1; fi; if [ $v = c5_noindex ]; then export CUTADAPT_B200_NO_INDEX_KERNEL=1; fi
  timeout 900 python bench.py --config 5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2v_bench_$v.json 2> gpurun_out/r2v_bench_$v.err
  show $v; tail -2 gpurun_out/r2v_bench_$v.err
done
unset CUTADAPT_B200_NO_LIGHT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:"cg_" -c 40 --csv --log-file gpurun_out/r2v_launches_c5.csv python bench.py --config 5 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2v_ncu_list.log 2>&1
grep "cg_" gpurun_out/r2v_launches_c5.csv | tail -12 | cut -c1-300
