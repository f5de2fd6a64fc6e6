#!/bin/bash
# round-1 session-2 sweep: warp-stream kernel with 256-bit value loads; EPL 4/8; CTAs per SM
mkdir -p gpurun_out
{
for epl in 4 8; do for w in 0 3; do
SUBLINEAR_B200_WARP_EPL=$epl SUBLINEAR_B200_WARP_CTAS=$w python tests/kernel_timing.py random 2>&1 | tail -1
done; done
python tests/kernel_timing.py banded 2>&1 | tail -1
python tests/kernel_timing.py random 1000000 2>&1 | tail -1
} > gpurun_out/s2_sweep.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
ncu --set full --clock-control none --import-source on -k regex:warp_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/s2_push python tests/kernel_timing.py random > gpurun_out/s2_ncu.log 2>&1
tail -3 gpurun_out/s2_sweep.log gpurun_out/s2_pytest.log; cat gpurun_out/s2_sweep.log; cat gpurun_out/s2_bench.json
