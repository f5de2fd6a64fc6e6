#!/bin/bash
# round 2 (N GPUs, N = $1): multi-rank tests for that world size, bench line, per-iteration timing of the bare recurrence
N=${1:-2}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_dist.py -m gpu -q -x -k "[$N-" > gpurun_out/r2_dist${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2_dist${N}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cut -c1-400 gpurun_out/r2_bench_n${N}.json; grep -o '"parity": {[^}]*}' gpurun_out/r2_bench_n${N}.json; grep -o '"roofline": {[^}]*}' gpurun_out/r2_bench_n${N}.json | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_timing.py 2>&1 | grep "^world" | tee gpurun_out/r2_dist${N}_timing.log
SUBLINEAR_B200_SLABS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_timing.py 2>&1 | grep "^world" | sed 's/^/SLABS=0 /' | tee -a gpurun_out/r2_dist${N}_timing.log
SUBLINEAR_B200_DEBUG_NOSTORE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/dist_timing.py 2>&1 | grep "^world" | tee -a gpurun_out/r2_dist${N}_timing.log
