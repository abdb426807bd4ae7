#!/bin/bash
# compute-sanitizer memcheck over the tests of the kernels added in round 2 (FASTQ: revcomp, pair adapters, paired
# demultiplexing, info rows; index kernel; fused statistics; specialised first stage)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests/test_gpu_fastq.py -q -m gpu -x -k "revcomp or info or pair_adapters or paired_demultiplexing or demultiplex_reference" > gpurun_out/r2h_memcheck_fastq.log 2>&1
echo "memcheck fastq rc=$?" >> gpurun_out/r2h_memcheck_fastq.log; tail -6 gpurun_out/r2h_memcheck_fastq.log
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "config5 or statistics_fused or specialised or bitplane" > gpurun_out/r2h_memcheck_parity.log 2>&1
echo "memcheck parity rc=$?" >> gpurun_out/r2h_memcheck_parity.log; tail -6 gpurun_out/r2h_memcheck_parity.log
