#!/bin/bash
mkdir -p gpurun_out
./sublinear-time-solver_b200/bench/numa_probe gpurun_out/numa_probe.bin 16 2>&1 | tail -2
python tests/cg_timing.py 10000000 5 2>&1 | tail -3 | tee gpurun_out/s7_cg_timing.log
python bench.py --workload banded_n10M_nnz100M --no-cpu-baseline --steps 5 > gpurun_out/s7_bench_banded.json 2>gpurun_out/s7.err; cut -c1-1200 gpurun_out/s7_bench_banded.json
python bench.py --workload c2_n1M_nnz10M --no-cpu-baseline --steps 10 > gpurun_out/s7_bench_c2.json 2>>gpurun_out/s7.err; cut -c1-1200 gpurun_out/s7_bench_c2.json
