#!/bin/bash
# round 2, 8 GPUs: oracle tests in the default (peer-memory) flavour — small system + slab-layout size —, the bench line, and
# the bare-recurrence timing in three layouts of the row blocks (default single pass, 2 slabs, 3 slabs)
N=8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -q -x -k "[8-p2p" > gpurun_out/r2_dist8_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_dist8_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; cut -c1-330 gpurun_out/r2_bench_n8.json; grep -o '"bit_exact": [a-z]*, "counts_equal": [a-z]*' gpurun_out/r2_bench_n8.json; grep -o '"avg_launch_us": [0-9.]*' gpurun_out/r2_bench_n8.json
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tests/dist_timing.py 2>&1 | grep "^world" | sed "s/^/$2 /" | tee -a gpurun_out/r2_dist8_timing.log; }
rm -f gpurun_out/r2_dist8_timing.log
run 29512 default
SUBLINEAR_B200_SLABS=2 run 29513 slabs2
SUBLINEAR_B200_SLABS=3 run 29514 slabs3
