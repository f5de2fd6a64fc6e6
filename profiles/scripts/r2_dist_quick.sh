#!/bin/bash
# round 2 (N GPUs, N = $1): bench line + per-iteration timing of the bare recurrence, no tests
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cut -c1-330 gpurun_out/r2_bench_n${N}.json; grep -o '"parity": {[^}]*}' gpurun_out/r2_bench_n${N}.json | cut -c100-330; grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r2_bench_n${N}.json
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tests/dist_timing.py 2>&1 | grep "^world" | sed "s/^/$2 /" | tee -a gpurun_out/r2_dist${N}_timing.log; }
rm -f gpurun_out/r2_dist${N}_timing.log
run 29512 default
SUBLINEAR_B200_SLAB_MIN_DENSITY=2 run 29513 slabs
SUBLINEAR_B200_DEBUG_NOSTORE=1 run 29514 default
SUBLINEAR_B200_SLAB_MIN_DENSITY=2 SUBLINEAR_B200_DEBUG_NOSTORE=1 run 29515 slabs
