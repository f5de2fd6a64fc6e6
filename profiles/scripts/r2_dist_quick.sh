#!/bin/bash
# round 2 (N GPUs, N = $1): bench line + per-iteration timing of the bare recurrence, no tests
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cut -c1-330 gpurun_out/r2_bench_n${N}.json; grep -o '"parity": {[^}]*}' gpurun_out/r2_bench_n${N}.json | cut -c100-330; grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r2_bench_n${N}.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_timing.py 2>&1 | grep "^world" | tee gpurun_out/r2_dist${N}_timing.log
SUBLINEAR_B200_SLABS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_timing.py 2>&1 | grep "^world" | sed 's/^/SLABS=0 /' | tee -a gpurun_out/r2_dist${N}_timing.log
SUBLINEAR_B200_DEBUG_NOSTORE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/dist_timing.py 2>&1 | grep "^world" | tee -a gpurun_out/r2_dist${N}_timing.log
SUBLINEAR_B200_SLABS=0 SUBLINEAR_B200_DEBUG_NOSTORE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tests/dist_timing.py 2>&1 | grep "^world" | sed 's/^/SLABS=0 /' | tee -a gpurun_out/r2_dist${N}_timing.log
