#!/bin/bash
mkdir -p gpurun_out
SUBLINEAR_B200_SLABS=2 timeout 300 python -m pytest tests/test_dist.py -m gpu -q > gpurun_out/d3_pytest_slabs.log 2>&1; echo "pytest(slabs=2) rc=$?"; grep -v "^Extension" gpurun_out/d3_pytest_slabs.log | tail -n 3
SUBLINEAR_B200_SLABS=3 SUBLINEAR_B200_DIST=nccl timeout 300 python -m pytest tests/test_dist.py -m gpu -q > gpurun_out/d3_pytest_slabs_nccl.log 2>&1; echo "pytest(slabs=3,nccl) rc=$?"; grep -v "^Extension" gpurun_out/d3_pytest_slabs_nccl.log | tail -n 3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/d3_bench.json 2> gpurun_out/d3_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/d3_bench.json; python -c "
import json; d=json.load(open('gpurun_out/d3_bench.json')); print(d['config']['device_layout'], d['roofline']['avg_launch_us'], d['config']['converged'], d['e2e']['value'])"
