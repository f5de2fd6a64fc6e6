#!/bin/bash
mkdir -p gpurun_out
python tests/slab_timing.py 10000000 10000000 5000000 3333333 2500000 1250000 2>&1 | grep columns | tee gpurun_out/slab_timing.log
