#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/f1_pytest.log | tail -n 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
{
SUBLINEAR_B200_WARP_CTAS=3 python tests/kernel_timing.py random 2>&1 | tail -1
SUBLINEAR_B200_WARP_CTAS=4 python tests/kernel_timing.py random 2>&1 | tail -1
SUBLINEAR_B200_WARP_EPL=8 SUBLINEAR_B200_SLABS=2 python tests/kernel_timing.py random 2>&1 | tail -1
python tests/kernel_timing.py banded 2>&1 | tail -1
} | tee gpurun_out/f1_timing.log
python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; cut -c1-200 gpurun_out/f1_bench.json
ncu --set full --clock-control none --import-source on -k regex:warp_kernel --launch-skip 13 --launch-count 3 -f -o gpurun_out/f1_slab_push python tests/kernel_timing.py random > gpurun_out/f1_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f1_bench_under_ncu.log 2>&1
tail -n 2 gpurun_out/f1_ncu.log
