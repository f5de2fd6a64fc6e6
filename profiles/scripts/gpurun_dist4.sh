#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/d4_bench.out 2> gpurun_out/d4_bench.err; echo "bench rc=$?"
echo "stdout lines: $(wc -l < gpurun_out/d4_bench.out)"; head -c 120 gpurun_out/d4_bench.out; echo; grep -c "NCCL version" gpurun_out/d4_bench.err
