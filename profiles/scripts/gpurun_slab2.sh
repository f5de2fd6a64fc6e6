#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layouts.py -m gpu -q -x > gpurun_out/sl_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/sl_pytest.log | tail -n 12
{
python tests/kernel_timing.py random 2>&1 | tail -1
SUBLINEAR_B200_SLABS=2 python tests/kernel_timing.py random 2>&1 | tail -1
SUBLINEAR_B200_SLABS=4 python tests/kernel_timing.py random 2>&1 | tail -1
} | tee gpurun_out/sl_timing.log
