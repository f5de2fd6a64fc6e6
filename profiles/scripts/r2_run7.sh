#!/bin/bash
# round 2, run 7 (1 GPU): full suite after the reduction-tail / exchange changes; C2 and headline timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2g_pytest.log
{
timeout 300 python tests/kernel_timing.py random 1000000 2>&1 | tail -1
timeout 300 python tests/kernel_timing.py random 2>&1 | tail -1
timeout 300 python tests/kernel_timing.py banded 2>&1 | tail -1
} | tee gpurun_out/r2g_timing.log
