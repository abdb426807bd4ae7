#!/bin/bash
# Round-2 verification after the fused statistics / pipe-balance changes: whole GPU suite, default bench line,
# A/B runs, other configurations, sanitizer, source-level captures of the list kernels, launch list with DRAM bytes.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -15 gpurun_out/r2p_pytest.log
export CUTADAPT_B200_STAGE_TIMES=1
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"; tail -c 2800 gpurun_out/r2p_bench.json; tail -5 gpurun_out/r2p_bench.err
show() {
python -c "
import json,sys
d = json.loads(open('gpurun_out/r2p_bench_$1.json').read().strip().split('\n')[-1]); print('$1: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f, kernel ms %.3f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']), d['roofline']['stage_ms_per_launch'], d['parity_mismatches'], d['config']['first_stage_specialised'])"
}
for v in imad0 imad15 nofuse; do
  unset CUTADAPT_B200_JIT_DEFINES CUTADAPT_B200_NO_FUSED_STATS
  case $v in
    imad0) export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0";;
    imad15) export CUTADAPT_B200_JIT_DEFINES="-DCG_CHAIN_IMAD=0x15";;
    nofuse) export CUTADAPT_B200_NO_FUSED_STATS=1;;
  esac
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2p_bench_$v.json 2> gpurun_out/r2p_bench_$v.err
  show $v
done
unset CUTADAPT_B200_JIT_DEFINES CUTADAPT_B200_NO_FUSED_STATS
for cfg in 3 4 5; do
  timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-e2e > gpurun_out/r2p_bench_c$cfg.json 2> gpurun_out/r2p_bench_c$cfg.err
  echo "config $cfg rc=$?"; show c$cfg; tail -2 gpurun_out/r2p_bench_c$cfg.err
done
unset CUTADAPT_B200_STAGE_TIMES
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2p_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2p_sanitizer_$tool.log
  tail -3 gpurun_out/r2p_sanitizer_$tool.log
done
export CUTADAPT_B200_JIT=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:"cg_" -c 120 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --reads 16000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2p_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)1" -s 2 -c 1 -o gpurun_out/r2p_plan -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2p_ncu_plan.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_list_kernel<\(bool\)0" -s 8 -c 1 -o gpurun_out/r2p_run -f python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2p_ncu_run.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_pscan" -s 2 -c 1 -o gpurun_out/r2p_pscan -f python bench.py --reads 16000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2p_ncu_pscan.log 2>&1
ls -la gpurun_out/r2p*
