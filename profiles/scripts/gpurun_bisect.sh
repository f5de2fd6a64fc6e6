#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_state_streaming.py tests/test_gpu_parity.py tests/test_io_formats.py tests/test_abi_cpu.py tests/test_dist.py tests/test_gpu_layouts.py tests/test_cg.py; do
  b=$(basename $f .py)
  MALLOC_CHECK_=3 timeout 600 python -X faulthandler -m pytest $f -m gpu -q -x > gpurun_out/bis_$b.log 2>&1
  echo "$b rc=$? $(grep -E 'passed|failed|error' gpurun_out/bis_$b.log | tail -1) $(grep -c -E 'malloc|free\(\)|corrupt|Aborted' gpurun_out/bis_$b.log)"
done
