#!/bin/bash
# FASTQ device paths (revcomp, pair adapters, paired demux, info rows), sanitizer runs of the smoke shape, bench lines of
# the other BASELINE configurations
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fastq.py -q -m gpu -x > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -30 gpurun_out/r2n_pytest.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2n_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2n_sanitizer_$tool.log
  tail -4 gpurun_out/r2n_sanitizer_$tool.log
done
export CUTADAPT_B200_STAGE_TIMES=1
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-e2e > gpurun_out/r2n_bench_c$cfg.json 2> gpurun_out/r2n_bench_c$cfg.err
  echo "config $cfg rc=$?"; tail -c 1500 gpurun_out/r2n_bench_c$cfg.json; tail -3 gpurun_out/r2n_bench_c$cfg.err
done
