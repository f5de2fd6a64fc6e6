#!/bin/bash
# round 2 (N GPUs, N = $1): multi-rank oracle tests for that world size (both exchange flavours; small system + a size whose
# row blocks take the column-slab layout), the bench line, per-iteration timing of the bare recurrence
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dist.py -m gpu -q -x -k "[$N-" > gpurun_out/r2_dist${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_dist${N}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cut -c1-330 gpurun_out/r2_bench_n${N}.json; grep -o '"bit_exact": [a-z]*, "counts_equal": [a-z]*' gpurun_out/r2_bench_n${N}.json; grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r2_bench_n${N}.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_timing.py 2>&1 | grep "^world" | tee gpurun_out/r2_dist${N}_timing.log
