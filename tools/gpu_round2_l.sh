#!/bin/bash
# Round-2 verification: the whole GPU suite, the default bench line (e2e + cpu_baseline), launch list with DRAM bytes,
# one --set full capture of the specialised first stage.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -6 gpurun_out/r2l_pytest.log
export CUTADAPT_B200_STAGE_TIMES=1
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2l_bench.json
tail -5 gpurun_out/r2l_bench.err
unset CUTADAPT_B200_STAGE_TIMES
export CUTADAPT_B200_JIT=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -c 400 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --reads 16000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2l_ncu_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_pscan" -s 2 -c 1 -o gpurun_out/r2l_pscan -f python bench.py --reads 16000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2l_ncu_pscan.log 2>&1
ls -la gpurun_out/r2l*
