#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layouts.py tests/test_cg.py -m gpu -x -q > gpurun_out/s4_pytest_new.log 2>&1
tail -n 5 gpurun_out/s4_pytest_new.log
{
for u in 8 4 5 10; do
SUBLINEAR_B200_SELL_U=$u python tests/kernel_timing.py random 2>&1 | tail -1
done
SUBLINEAR_B200_SELL_U=8 SUBLINEAR_B200_SELL_CTAS=3 python tests/kernel_timing.py random 2>&1 | tail -1
SUBLINEAR_B200_SELL_U=4 SUBLINEAR_B200_SELL_CTAS=4 python tests/kernel_timing.py random 2>&1 | tail -1
python tests/kernel_timing.py banded 2>&1 | tail -1
python tests/kernel_timing.py random 1000000 2>&1 | tail -1
SUBLINEAR_B200_SELL=0 python tests/kernel_timing.py random 2>&1 | tail -1
} > gpurun_out/s4_sweep.log 2>&1
cat gpurun_out/s4_sweep.log
