#!/bin/bash
# Re-run of the final evidence after the last kernel changes (index kernel, 40-row run kernel, cooperative quality scan)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -6 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
export CUTADAPT_B200_STAGE_TIMES=1
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2g_bench.json; tail -5 gpurun_out/r2g_bench.err
for cfg in 4 5; do
  timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r2g_bench_c$cfg.json 2> gpurun_out/r2g_bench_c$cfg.err
  echo "config $cfg rc=$?"; tail -c 400 gpurun_out/r2g_bench_c$cfg.json; tail -2 gpurun_out/r2g_bench_c$cfg.err
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2g_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2g_sanitizer_$tool.log
  tail -3 gpurun_out/r2g_sanitizer_$tool.log
done
