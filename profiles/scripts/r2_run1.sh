#!/bin/bash
# round 2, run 1 (1 GPU): layout + full-size parity tests on the fused slab kernel, slab sweeps, first ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_layouts.py tests/test_full_size_parity.py -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/r2a_pytest.log | tail -n 15
{
timeout 600 python tests/slab_sweep.py 10000000 full blocks 2>&1 | grep -E "^(full|block)"
SUBLINEAR_B200_SLAB_CTAS=4 timeout 600 python tests/slab_sweep.py 10000000 full 2>&1 | grep -E "^(full|block)"
} | tee gpurun_out/r2a_sweep.log
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:slab_kernel --launch-skip 14 --launch-count 1 -f -o gpurun_out/r2a_slab_push python tests/kernel_timing.py random > gpurun_out/r2a_ncu.log 2>&1
tail -n 3 gpurun_out/r2a_ncu.log
