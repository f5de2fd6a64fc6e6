#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/f3_pytest.log | tail -n 25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
