#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -q > gpurun_out/d2_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^Extension" gpurun_out/d2_pytest.log | tail -n 5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/d2_bench.json 2> gpurun_out/d2_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/d2_bench.json; tail -n 3 gpurun_out/d2_bench.err
