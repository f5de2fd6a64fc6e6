#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layouts.py -m gpu -q -x > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/g_pytest.log | tail -n 2
python tests/kernel_timing.py random 2>&1 | tail -1 | tee gpurun_out/g_timing.log
timeout 200 python tests/config_timing.py c4 2>&1 | grep config | tee gpurun_out/g_c4.json
timeout 400 python tests/config_timing.py c3 2>&1 | grep config | tee gpurun_out/g_c3.json
