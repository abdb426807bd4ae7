#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2last2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2last2_pytest.log; tail -6 gpurun_out/r2last2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -c 700
