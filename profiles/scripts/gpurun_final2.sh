#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/f2_pytest.log | tail -n 4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 150 python tests/config_timing.py c3 2>&1 | grep config | tee gpurun_out/f2_c3.json
timeout 200 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; cut -c1-180 gpurun_out/f2_bench.json
