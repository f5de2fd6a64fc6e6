#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s6_pytest.log 2>&1
echo "pytest rc=$?"
grep -v "^Extension modules" gpurun_out/s6_pytest.log | tail -n 6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
