#!/bin/bash
# round 2, run 3 (1 GPU): slab kernel v2 (row offsets instead of lengths, uniform accumulate loop), variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layouts.py -m gpu -x -q 2>&1 | tail -n 2
{
for v in 0 1 2 3; do
SUBLINEAR_B200_SLAB_VARIANT=$v timeout 600 python tests/kernel_timing.py random 2>&1 | tail -1
done
SUBLINEAR_B200_SLAB_VARIANT=1 timeout 600 python tests/slab_sweep.py 10000000 blocks 2>&1 | grep -E "^(full|block)"
} | tee gpurun_out/r2c_sweep.log
for v in 0 1; do
SUBLINEAR_B200_SLAB_VARIANT=$v timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:slab_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/r2c_slab_push_v$v python tests/kernel_timing.py random > gpurun_out/r2c_ncu_v$v.log 2>&1
done
tail -n 2 gpurun_out/r2c_ncu_v1.log
