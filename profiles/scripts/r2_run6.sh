#!/bin/bash
# round 2, run 6 (1 GPU): full GPU suite + smoke + bench line on the fused slab kernel
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/r2f_pytest.log | tail -n 8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; cut -c1-1500 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
