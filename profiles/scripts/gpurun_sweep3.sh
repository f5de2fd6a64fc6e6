#!/bin/bash
mkdir -p gpurun_out
{
for pr in 0 1 2 3; do
SUBLINEAR_B200_WARP_PROBE=$pr python tests/kernel_timing.py random 2>&1 | tail -1
done
SUBLINEAR_B200_WARP_PROBE=2 SUBLINEAR_B200_WARP_CTAS=3 python tests/kernel_timing.py random 2>&1 | tail -1
SUBLINEAR_B200_WARP_PROBE=1 SUBLINEAR_B200_WARP_CTAS=3 python tests/kernel_timing.py random 2>&1 | tail -1
} > gpurun_out/s3_sweep.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest.log 2>&1
cat gpurun_out/s3_sweep.log; tail -n 15 gpurun_out/s3_pytest.log
