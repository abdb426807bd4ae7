#!/bin/bash
# Final round-2 evidence: whole GPU suite, bench lines (default with e2e + cpu_baseline, reference arm, configs 3-5),
# launch list with DRAM bytes, --set full of the three kernels of the pass
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
export CUTADAPT_B200_STAGE_TIMES=1
timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
echo "reference rc=$?"; tail -c 600 gpurun_out/r2f_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/r2f_bench_c$cfg.json 2> gpurun_out/r2f_bench_c$cfg.err
  echo "config $cfg rc=$?"; tail -c 1200 gpurun_out/r2f_bench_c$cfg.json; tail -2 gpurun_out/r2f_bench_c$cfg.err
done
unset CUTADAPT_B200_STAGE_TIMES
export CUTADAPT_B200_JIT=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:"cg_" -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --reads 16000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_pscan" -s 2 -c 1 -o gpurun_out/r2f_pscan -f python bench.py --reads 16000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu_pscan.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)1" -s 2 -c 1 -o gpurun_out/r2f_plan -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu_plan.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)0" -s 8 -c 1 -o gpurun_out/r2f_run -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_ncu_run.log 2>&1
ls -la gpurun_out/r2f*
