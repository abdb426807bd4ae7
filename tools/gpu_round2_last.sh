#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_z_more_goldens.py -q -m gpu -x -k "rest_and_wildcard or info or paired_revcomp or revcomp_and_pair" > gpurun_out/r2last_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2last_pytest.log; tail -15 gpurun_out/r2last_pytest.log
