#!/bin/bash
# round 2, run 4 (1 GPU): slab kernel v3 (simpler walk, padded product staging, late stream prefetch, hint on/off)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layouts.py -m gpu -x -q 2>&1 | tail -n 2
{
for v in 0 1 2 3 4 5; do
SUBLINEAR_B200_SLAB_VARIANT=$v timeout 600 python tests/kernel_timing.py random 2>&1 | tail -1
done
} | tee gpurun_out/r2d_sweep.log
for v in 1 2 5; do
SUBLINEAR_B200_SLAB_VARIANT=$v timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sectors_srcunit_ltcfabric.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max --clock-control none --cache-control none -k regex:slab_kernel --launch-skip 14 --launch-count 1 python tests/kernel_timing.py random 2>&1 | grep -E "slab_kernel|inst_executed|duration|lts__|dram__|l1tex|sm__cycles" | tee -a gpurun_out/r2d_ncu_metrics.log
done
