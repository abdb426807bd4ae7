#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"cg_index_kernel" -s 2 -c 1 -o gpurun_out/r2_index -f python bench.py --config 5 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_ncu_index.log 2>&1
ls -la gpurun_out/r2_index.ncu-rep
