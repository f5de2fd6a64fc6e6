#!/bin/bash
# quick validation of the exchange protocol at N = $1: small oracle test in both flavours + bare-recurrence timing
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dist.py -m gpu -q -x -k "test_dist_solve_vs_oracle and [$N-" 2>&1 | tail -n 4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_timing.py 2>&1 | grep "^world" | tee gpurun_out/r2_dist${N}_timing.log
