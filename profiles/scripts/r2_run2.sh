#!/bin/bash
# round 2, run 2 (1 GPU): fused slab kernel after forcing the four gathers of a lane into flight together
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layouts.py -m gpu -x -q 2>&1 | tail -n 2
{
timeout 600 python tests/slab_sweep.py 10000000 full blocks 2>&1 | grep -E "^(full|block)"
} | tee gpurun_out/r2b_sweep.log
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:slab_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/r2b_slab_push python tests/kernel_timing.py random > gpurun_out/r2b_ncu.log 2>&1
tail -n 2 gpurun_out/r2b_ncu.log
