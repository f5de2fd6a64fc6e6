#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --workload c5_n100M_nnz1B > gpurun_out/c5_bench_n8.json 2> gpurun_out/c5_bench_n8.err; echo "c5 rc=$?"
cut -c1-2500 gpurun_out/c5_bench_n8.json; grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/c5_bench_n8.err | tail -n 5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c3_bench_n8_sell.json 2> gpurun_out/c3_bench_n8.err; echo "c3x8 rc=$?"
cut -c1-600 gpurun_out/c3_bench_n8_sell.json
